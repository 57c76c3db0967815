"""Seeded synthetic inputs and reference-format random ``state_dict``s.

Host-side NumPy/torch-CPU generators only: there are no datasets or checkpoints in
this environment, so tests, ``bench.py`` and ``__graft_entry__.smoke()`` all draw
their inputs and weights from here.  Key names / shapes follow the reference's
modules so the same dict loads into the reference (oracle side) and into the
drop-in modules of this package:
  video  -- models/video_models/model.py:81-85, resnet.py:28-69, 99-118
  audio  -- models/audio_models/tdnn.py:25-34, 59-87
  fusion -- models/fusion_models/model_fusion.py:11-17
"""
import math
import numpy as np
import torch

SEED = 1  # the only seed the reference uses (train_video.py:70-73)

ETDNN_OPTS = {
    'arch': 'etdnn',
    'tdnn': dict(input_dim=24, hidden_dim=[512, 512, 512, 512, 1500],
                 context=[[-2, -1, 0, 1, 2], [-2, 0, 2], [-3, 0, 3], [0], [0]],
                 tdnn_layers=5, fc_layers=3, embedding_dim=512, pooling='statistic',
                 attention_hidden_size=64, bn_first=True),
    'etdnn': dict(input_dim=24, hidden_dim=[512] * 9 + [1500],
                  context=[[-2, -1, 0, 1, 2], [0], [-2, 0, 2], [0], [-3, 0, 3], [0], [-4, 0, 4], [0], [0], [0]],
                  tdnn_layers=10, fc_layers=3, embedding_dim=512, pooling='statistic',
                  attention_hidden_size=64, bn_first=True),
}   # conf/fusion_config.yaml:48-70

TCN_OPTIONS = dict(num_layers=4, kernel_size=[3, 5, 7], dropout=0.2, dwpw=False, width_mult=1)


def audio_opts(arch='etdnn', pooling='statistic'):
    import copy
    o = copy.deepcopy(ETDNN_OPTS)
    o['arch'] = arch
    o[arch]['pooling'] = pooling
    return o


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _bn(sd, p, c, rng, randomize):
    if randomize:   # exercise folding / epilogues
        sd[p + '.weight'] = _t(rng.uniform(0.5, 1.5, c))
        sd[p + '.bias'] = _t(rng.normal(0, 0.2, c))
        sd[p + '.running_mean'] = _t(rng.normal(0, 0.2, c))
        sd[p + '.running_var'] = _t(rng.uniform(0.5, 1.5, c))
    else:
        sd[p + '.weight'] = torch.ones(c)
        sd[p + '.bias'] = torch.zeros(c)
        sd[p + '.running_mean'] = torch.zeros(c)
        sd[p + '.running_var'] = torch.ones(c)
    sd[p + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)


def _prelu(sd, key, c, rng, randomize):
    sd[key] = _t(rng.uniform(0.0, 0.5, c)) if randomize else torch.full((c,), 0.25)


def make_video_state_dict(seed=SEED, randomize=True, relu_type='prelu'):
    """Lipreading(backbone 'resnet') keys without the (unused at extract time) TCN head."""
    rng = np.random.default_rng(seed)
    sd = {}
    fan_in = 5 * 7 * 7
    sd['frontend3D.0.weight'] = _t(rng.uniform(-1, 1, (64, 1, 5, 7, 7)) / math.sqrt(fan_in))
    _bn(sd, 'frontend3D.1', 64, rng, randomize)
    if relu_type == 'prelu':
        _prelu(sd, 'frontend3D.2.weight', 64, rng, randomize)
    inpl = 64
    for li, planes in enumerate([64, 128, 256, 512]):
        for bi in range(2):
            p = 'trunk.layer%d.%d.' % (li + 1, bi)
            stride = 2 if (li > 0 and bi == 0) else 1
            cin = inpl if bi == 0 else planes
            sd[p + 'conv1.weight'] = _t(rng.normal(0, math.sqrt(2.0 / (9 * planes)), (planes, cin, 3, 3)))
            _bn(sd, p + 'bn1', planes, rng, randomize)
            sd[p + 'conv2.weight'] = _t(rng.normal(0, math.sqrt(2.0 / (9 * planes)), (planes, planes, 3, 3)))
            _bn(sd, p + 'bn2', planes, rng, randomize)
            if relu_type == 'prelu':
                _prelu(sd, p + 'relu1.weight', planes, rng, randomize)
                _prelu(sd, p + 'relu2.weight', planes, rng, randomize)
            if bi == 0 and (stride != 1 or cin != planes):
                sd[p + 'downsample.0.weight'] = _t(rng.normal(0, math.sqrt(2.0 / planes), (planes, cin, 1, 1)))
                _bn(sd, p + 'downsample.1', planes, rng, randomize)
        inpl = planes
    return sd


def make_tcn_state_dict(num_classes=500, hidden=256, ksizes=(3, 5, 7), layers=4, in_dim=512, seed=SEED,
                        randomize=True):
    """MS-TCN head keys of Lipreading (models/video_models/tcn.py:62-140, model.py:20-37):
    tcn.mb_ms_tcn.network.{i}.cbcr{0,1}_{k}.{conv,batchnorm,non_lin}, .downsample, .relu_final, tcn.tcn_output."""
    rng = np.random.default_rng(seed + 3000)
    sd = {}
    nk = len(ksizes)
    width = hidden * nk
    for i in range(layers):
        cin = in_dim if i == 0 else width
        p = 'tcn.mb_ms_tcn.network.%d.' % i
        for stage, c_in in ((0, cin), (1, width)):
            for k_idx, k in enumerate(ksizes):
                q = p + 'cbcr%d_%d.' % (stage, k_idx)
                bound = 1.0 / math.sqrt(c_in * k)
                sd[q + 'conv.weight'] = _t(rng.uniform(-bound, bound, (hidden, c_in, k)) * math.sqrt(3.0))
                sd[q + 'conv.bias'] = _t(rng.uniform(-bound, bound, hidden))
                _bn(sd, q + 'batchnorm', hidden, rng, randomize)
                _prelu(sd, q + 'non_lin.weight', hidden, rng, randomize)
        if cin // nk != width:
            bound = 1.0 / math.sqrt(cin)
            sd[p + 'downsample.weight'] = _t(rng.uniform(-bound, bound, (width, cin, 1)))
            sd[p + 'downsample.bias'] = _t(rng.uniform(-bound, bound, width))
        _prelu(sd, p + 'relu_final.weight', width, rng, randomize)
    bound = 1.0 / math.sqrt(width)
    sd['tcn.tcn_output.weight'] = _t(rng.uniform(-bound, bound, (num_classes, width)))
    sd['tcn.tcn_output.bias'] = _t(rng.uniform(-bound, bound, num_classes))
    return sd


def make_audio_state_dict(opts, seed=SEED, randomize=True):
    rng = np.random.default_rng(seed + 1000)
    o = opts[opts['arch']]
    sd = {}
    cin = o['input_dim']
    for i in range(o['tdnn_layers']):
        k = len(o['context'][i])
        cout = o['hidden_dim'][i]
        bound = 1.0 / math.sqrt(cin * k)
        p = 'tdnn.%d.' % i
        sd[p + 'context_layer.weight'] = _t(rng.uniform(-bound, bound, (cout, cin, k)) * math.sqrt(3.0))
        sd[p + 'context_layer.bias'] = _t(rng.uniform(-bound, bound, cout))
        _bn(sd, p + 'bn', cout, rng, randomize)
        cin = cout
    pool = o['pooling']
    pooled = cin * 2 if pool in ('statistic', 'attentive_statistic') else cin
    if pool == 'attentive_statistic':
        h = o['attention_hidden_size']
        sd['pooling.W'] = _t(rng.normal(0, math.sqrt(2.0 / (h + cin)), (h, cin)))
        sd['pooling.b'] = _t(rng.normal(0, math.sqrt(2.0 / (1 + h)), (1, h)))
        sd['pooling.v'] = _t(rng.normal(0, math.sqrt(2.0 / (1 + h)), (h, 1)))
        sd['pooling.k'] = _t(rng.normal(0, 1.0, (1, 1)))
    e = o['embedding_dim']
    b1 = 1.0 / math.sqrt(pooled)
    sd['fc1.weight'] = _t(rng.uniform(-b1, b1, (e, pooled)))
    sd['fc1.bias'] = _t(rng.uniform(-b1, b1, e))
    _bn(sd, 'bn1', e, rng, randomize)
    b2 = 1.0 / math.sqrt(e)
    sd['fc2.weight'] = _t(rng.uniform(-b2, b2, (e, e)))
    sd['fc2.bias'] = _t(rng.uniform(-b2, b2, e))
    _bn(sd, 'bn2', e, rng, randomize)
    return sd


AUDIO_RESNET_OPTS = {'arch': 'resnet',
                     'resnet': dict(input_dim=1, hidden_dim=[64, 128, 256], residual_block_layers=[3, 3, 3],
                                    fc_layers=1, embedding_dim=256, pooling='average')}   # conf/audio_config.yaml:93-102


def make_audio_resnet_state_dict(opts=None, seed=SEED, randomize=True):
    """Weights for the build-defined audio ResNet (deeplip_b200/audio_models/resnet.py)."""
    o = (opts or AUDIO_RESNET_OPTS)['resnet']
    rng = np.random.default_rng(seed + 4000)
    sd = {}
    hidden, blocks = o['hidden_dim'], o['residual_block_layers']
    sd['conv1.weight'] = _t(rng.normal(0, math.sqrt(2.0 / 9), (hidden[0], 1, 3, 3)))
    _bn(sd, 'bn0', hidden[0], rng, randomize)
    inpl = hidden[0]
    for i, (planes, nb) in enumerate(zip(hidden, blocks)):
        for b in range(nb):
            p = 'layer%d.%d.' % (i + 1, b)
            stride = 2 if (i > 0 and b == 0) else 1
            sd[p + 'conv1.weight'] = _t(rng.normal(0, math.sqrt(2.0 / (9 * planes)), (planes, inpl, 3, 3)))
            _bn(sd, p + 'bn1', planes, rng, randomize)
            sd[p + 'conv2.weight'] = _t(rng.normal(0, math.sqrt(2.0 / (9 * planes)), (planes, planes, 3, 3)))
            _bn(sd, p + 'bn2', planes, rng, randomize)
            if stride != 1 or inpl != planes:
                sd[p + 'downsample.0.weight'] = _t(rng.normal(0, math.sqrt(2.0 / planes), (planes, inpl, 1, 1)))
                _bn(sd, p + 'downsample.1', planes, rng, randomize)
            inpl = planes
    pooled = inpl * (2 if o.get('pooling', 'average') == 'statistic' else 1)
    e = o['embedding_dim']
    b1 = 1.0 / math.sqrt(pooled)
    sd['fc1.weight'] = _t(rng.uniform(-b1, b1, (e, pooled)))
    sd['fc1.bias'] = _t(rng.uniform(-b1, b1, e))
    _bn(sd, 'bn1', e, rng, randomize)
    return sd


def make_fusion_state_dict(input_size=1024, hidden=512, seed=SEED, randomize=True):
    rng = np.random.default_rng(seed + 2000)
    sd = {}
    b1 = 1.0 / math.sqrt(input_size)
    sd['fc1.weight'] = _t(rng.uniform(-b1, b1, (hidden, input_size)))
    sd['fc1.bias'] = _t(rng.uniform(-b1, b1, hidden))
    _bn(sd, 'bn1', hidden, rng, randomize)
    b2 = 1.0 / math.sqrt(hidden)
    sd['fc2.weight'] = _t(rng.uniform(-b2, b2, (hidden, hidden)))
    sd['fc2.bias'] = _t(rng.uniform(-b2, b2, hidden))
    return sd


# ----------------------------------------------------------------------------- inputs
def _smooth_field(rng, h, w, cut=6):
    """Low-pass N(0,1) field, unit variance."""
    f = np.fft.rfft2(rng.standard_normal((h, w)))
    ky = np.minimum(np.arange(h), h - np.arange(h))[:, None]
    kx = np.arange(f.shape[1])[None, :]
    f *= np.exp(-(ky ** 2 + kx ** 2) / (2.0 * cut * cut))
    x = np.fft.irfft2(f, s=(h, w))
    return (x - x.mean()) / (x.std() + 1e-12)


def lip_crops_u8(speakers, T=75, H=96, W=96, seed=SEED, utt_sigma=0.35, frame_sigma=8.0):
    """(len(speakers), T, H, W) uint8 GRID-shaped lip crops with speaker structure:
    a fixed smooth field per speaker (mean 107, std 42 grey levels) + a weaker
    per-utterance field + per-frame noise sigma 8, clipped to [0,255]."""
    speakers = np.asarray(speakers)
    out = np.empty((len(speakers), T, H, W), dtype=np.uint8)
    rng_u = np.random.default_rng(seed + 77)
    fields = {}
    for i, s in enumerate(speakers):
        s = int(s)
        if s not in fields:
            fields[s] = _smooth_field(np.random.default_rng(seed * 100003 + s), H, W)
        base = fields[s] + utt_sigma * _smooth_field(rng_u, H, W)
        x = 107.0 + 42.0 * base[None] + frame_sigma * rng_u.standard_normal((T, H, W))
        out[i] = np.clip(np.rint(x), 0, 255).astype(np.uint8)
    return out


def speech_like_audio(speakers, nsamp=48000, rate=16000, seed=SEED, noise=0.1):
    """(len(speakers), nsamp) float32 in [-1,1]: per speaker a fixed harmonic stack
    (f0 ~ U(90,250) Hz, 20 partials) + per-utterance white noise."""
    speakers = np.asarray(speakers)
    out = np.empty((len(speakers), nsamp), dtype=np.float32)
    t = np.arange(nsamp) / rate
    rng_u = np.random.default_rng(seed + 99)
    for i, s in enumerate(speakers):
        r = np.random.default_rng(seed * 7919 + int(s))
        f0 = r.uniform(90, 250)
        amp = np.abs(r.standard_normal(20))
        ph = r.uniform(0, 2 * np.pi, 20)
        x = sum(a * np.sin(2 * np.pi * f0 * (k + 1) * t + p) for k, (a, p) in enumerate(zip(amp, ph)))
        x = x / (np.abs(x).max() + 1e-9) * 0.5 + noise * rng_u.standard_normal(nsamp)
        out[i] = np.clip(x, -1, 1)
    return out


def speaker_of_utt(utt):
    """'s14/prbv1p.wav' -> 14 ; 's39_l_lrwo3a.wav' -> 39 (trial id formats)."""
    head = utt.split('/')[0].split('_')[0]
    return int(head[1:])


def structured_embeddings(speakers, dim=1024, seed=SEED, within=0.6):
    """(N,dim) float32 embeddings = speaker centroid + within-speaker noise; used to
    exercise scoring at full trial-list size without running extraction."""
    speakers = np.asarray(speakers)
    rng = np.random.default_rng(seed + 5)
    cents = {}
    out = np.empty((len(speakers), dim), dtype=np.float32)
    for i, s in enumerate(speakers):
        s = int(s)
        if s not in cents:
            cents[s] = np.random.default_rng(seed * 31337 + s).standard_normal(dim)
        out[i] = cents[s] + within * rng.standard_normal(dim)
    return out


def make_trial_file(path, kind='grid', seed=1, n_target=4000, n_non=16000):
    """Synthetic trial list with the shape of database/trial_{grid,lomgrid}_v1.txt:
    targets first, then non-targets; '<label> <utt1> <utt2>' + trailing TAB (grid) / space (lomgrid)."""
    rng = np.random.default_rng(seed)
    spk = [s for s in range(1, 35) if s != 21] if kind == 'grid' else list(range(2, 56, 1))[:36]
    per = 900 if kind == 'grid' else 100
    def utt(s, i):
        return 's%d/u%05d.wav' % (s, i) if kind == 'grid' else 's%d_l_u%05d.wav' % (s, i)
    tail = '\t' if kind == 'grid' else ' '
    lines = []
    for _ in range(n_target):
        s = spk[rng.integers(len(spk))]
        a, b = rng.choice(per, 2, replace=False)
        lines.append('1 %s %s%s' % (utt(s, a), utt(s, b), tail))
    for _ in range(n_non):
        s1, s2 = rng.choice(len(spk), 2, replace=False)
        lines.append('0 %s %s%s' % (utt(spk[s1], rng.integers(per)), utt(spk[s2], rng.integers(per)), tail))
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    return path
