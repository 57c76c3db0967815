"""fp32 PyTorch (CPU) restatement of the reference's model forwards, driven by a
reference-format ``state_dict``.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

PINNED: ``tests/test_oracle.py`` checks every function here against the
reference's own modules imported from /root/reference (when present) and against
golden vectors under tests/golden/ that ``oracle/gen_golden.py`` produced by
running those modules.

Written functionally (no nn.Module mirrors) so the arithmetic order is explicit:
each function cites the reference lines it restates.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm{1,2,3}d default, used everywhere in the reference


def _strip(sd):
    """Tolerate the DataParallel 'module.' prefix (models/audio_models/tdnn.py:128)."""
    return {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd[p + '.weight'], sd[p + '.bias'], False, 0.0, BN_EPS)


# ----------------------------------------------------------------------------- audio
def tdnn_context_to_conv(context):
    """models/audio_models/tdnn.py:17-22: context list -> (kernel_size, dilation)."""
    k = len(context)
    d = (context[-1] - context[0]) // (k - 1) if k > 1 else 1
    return k, d


def mean_std_pooling(x):
    """models/audio_models/pooling.py:18-26: mean || unbiased std over time."""
    return torch.cat([x.mean(dim=2), x.std(dim=2)], dim=1)


def attentive_stat_pooling(sd, x, prefix='pooling.'):
    """models/audio_models/pooling.py:89-107 (no eps / clamp under the sqrt)."""
    W, b, v, k = (sd[prefix + n] for n in ('W', 'b', 'v', 'k'))
    h = torch.relu(W.matmul(x).transpose(1, 2) + b)      # (B,T,H)
    alpha = torch.softmax(h.matmul(v) + k, dim=1)         # (B,T,1)
    mean = torch.matmul(x, alpha).squeeze(-1)
    ex2 = torch.matmul(x * x, alpha).squeeze(-1)
    return torch.cat([mean, torch.sqrt(ex2 - mean * mean)], dim=1)


def tdnn_trunk(sd, x, arch_opts):
    """models/audio_models/tdnn.py:35-43, 59-63: valid dilated Conv1d -> BN -> LeakyReLU(0.2)
    (bn_first) or Conv1d -> LeakyReLU -> BN."""
    for i in range(arch_opts['tdnn_layers']):
        k, d = tdnn_context_to_conv(arch_opts['context'][i])
        p = 'tdnn.%d.' % i
        x = F.conv1d(x, sd[p + 'context_layer.weight'], sd[p + 'context_layer.bias'], dilation=d)
        if arch_opts['bn_first']:
            x = F.leaky_relu(_bn(x, sd, p + 'bn'), 0.2)
        else:
            x = _bn(F.leaky_relu(x, 0.2), sd, p + 'bn')
    return x


def speaker_extract_embedding(sd, x, opts):
    """models/audio_models/tdnn.py:89-101.  x: (B,F,T) -> (xv, x_a)."""
    sd = _strip(sd)
    o = opts[opts['arch']]
    h = tdnn_trunk(sd, x, o)
    if o['pooling'] == 'statistic':
        h = mean_std_pooling(h)
    elif o['pooling'] == 'average':
        # reference bug: AdaptiveAvgPool1d(1) leaves (B,C,1) and squeeze_(1) is a no-op, so
        # fc1 raises a shape error (tdnn.py:69, 91-93) -- unrunnable upstream, not restated.
        raise NotImplementedError("pooling 'average' fails in the reference (tdnn.py:91-93)")
    elif o['pooling'] == 'attentive_statistic':
        h = attentive_stat_pooling(sd, h)
    else:
        raise NotImplementedError('Other pooling method has not implemented.')
    x_a = F.linear(h, sd['fc1.weight'], sd['fc1.bias'])
    if o['bn_first']:
        h = F.leaky_relu(_bn(x_a, sd, 'bn1'), 0.2)
    else:
        h = _bn(F.leaky_relu(x_a, 0.2), sd, 'bn1')
    xv = F.linear(h, sd['fc2.weight'], sd['fc2.bias'])
    return xv, x_a


def speaker_forward(sd, x, opts):
    """models/audio_models/tdnn.py:103-111."""
    sd = _strip(sd)
    xv, _ = speaker_extract_embedding(sd, x, opts)
    if opts[opts['arch']]['bn_first']:
        return F.leaky_relu(_bn(xv, sd, 'bn2'), 0.2)
    return _bn(F.leaky_relu(xv, 0.2), sd, 'bn2')


def audio_resnet_extract_embedding(sd, x, opts):
    """fp32 restatement of deeplip_b200.audio_models.resnet.SpeakerEmbNet -- a BUILD-DEFINED model: the
    reference's `models/resnet.py` does not exist, only conf/audio_config.yaml:93-102 and the
    call sites train_audio.py:64-66, 183-184, 250-252.  PARITY UNPINNED with respect to the reference.
    x: (B,1,F,T) -> (xv, x_a)."""
    sd = _strip(sd)
    o = opts[opts['arch']] if 'arch' in opts else opts
    h = F.relu(_bn(F.conv2d(x, sd['conv1.weight'], None, padding=1), sd, 'bn0'))
    for i, nb in enumerate(o['residual_block_layers']):
        for b in range(nb):
            p = 'layer%d.%d.' % (i + 1, b)
            stride = 2 if (i > 0 and b == 0) else 1
            out = F.relu(_bn(F.conv2d(h, sd[p + 'conv1.weight'], None, stride=stride, padding=1), sd, p + 'bn1'))
            out = _bn(F.conv2d(out, sd[p + 'conv2.weight'], None, padding=1), sd, p + 'bn2')
            res = h
            if p + 'downsample.0.weight' in sd:
                res = _bn(F.conv2d(h, sd[p + 'downsample.0.weight'], None, stride=stride), sd, p + 'downsample.1')
            h = F.relu(out + res)
    flat = h.flatten(2)
    pooled = flat.mean(dim=2) if o.get('pooling', 'average') == 'average' else torch.cat([flat.mean(2), flat.std(2)], 1)
    xa = F.linear(pooled, sd['fc1.weight'], sd['fc1.bias'])
    return xa, xa


# ----------------------------------------------------------------------------- video
def video_preprocess(frames_u8, crop=88, mean=0.421, std=0.165):
    """models/video_models/dataloaders.py:19-24 + preprocess.py:60-68, 80-92.
    (T,H,W) uint8 -> (T,88,88) float32:  x/255 -> centre crop -> (x-mean)/std,
    computed in float64 like the NumPy pipeline, cast at the end
    (models/fusion_models/datasets.py:374)."""
    x = frames_u8.to(torch.float64)
    x = (x - 0.0) / 255.0
    t, h, w = x.shape
    dw = int(round((w - crop)) / 2.)
    dh = int(round((h - crop)) / 2.)
    x = x[:, dh:dh + crop, dw:dw + crop]
    return ((x - mean) / std).to(torch.float32)


def _act(x, sd, key):
    """nn.PReLU(num_parameters=C) when the key exists, else ReLU
    (models/video_models/resnet.py:39-45)."""
    return F.prelu(x, sd[key]) if key in sd else F.relu(x)


def video_frontend3d(sd, x):
    """models/video_models/model.py:81-85."""
    x = F.conv3d(x, sd['frontend3D.0.weight'], None, stride=(1, 2, 2), padding=(2, 3, 3))
    x = _act(_bn(x, sd, 'frontend3D.1'), sd, 'frontend3D.2.weight')
    return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))


def basic_block(sd, x, p, stride):
    """models/video_models/resnet.py:56-69."""
    out = F.conv2d(x, sd[p + 'conv1.weight'], None, stride=stride, padding=1)
    out = _act(_bn(out, sd, p + 'bn1'), sd, p + 'relu1.weight')
    out = F.conv2d(out, sd[p + 'conv2.weight'], None, stride=1, padding=1)
    out = _bn(out, sd, p + 'bn2')
    if p + 'downsample.0.weight' in sd:      # 1x1 stride-s conv + BN, resnet.py:13-17
        res = _bn(F.conv2d(x, sd[p + 'downsample.0.weight'], None, stride=stride), sd, p + 'downsample.1')
    else:
        res = x
    return _act(out + res, sd, p + 'relu2.weight')


def resnet_trunk(sd, x, layers=(2, 2, 2, 2)):
    """models/video_models/resnet.py:120-127 (+ AdaptiveAvgPool2d(1), view)."""
    for li, nb in enumerate(layers):
        for bi in range(nb):
            stride = 2 if (li > 0 and bi == 0) else 1
            x = basic_block(sd, x, 'trunk.layer%d.%d.' % (li + 1, bi), stride)
    return x.mean(dim=(2, 3))


def lipreading_features(sd, x):
    """models/video_models/model.py:96-105 with extract_feats=True.
    x: (B,1,T,H,W) f32 -> (B,T,512)."""
    sd = _strip(sd)
    B, C, T, H, W = x.shape
    x = video_frontend3d(sd, x)
    Tn = x.shape[2]
    x = x.transpose(1, 2).reshape(B * Tn, x.shape[1], x.shape[3], x.shape[4])   # model.py:9-13
    x = resnet_trunk(sd, x)
    return x.view(B, Tn, x.size(1))


def ms_tcn_logits(sd, feats, lengths, ksizes=(3, 5, 7), layers=4):
    """MS-TCN head, models/video_models/tcn.py:28-140 + model.py:20-37 (eval mode, dwpw=False):
    per block two sets of {Conv1d(k, dilation 2^i, padding (k-1)2^i) -> BN -> symmetric chomp -> PReLU}
    branches concatenated, + 1x1-conv skip, PReLU; then the masked temporal mean (`_average_batch`) and Linear.
    feats: (B,T,512) -> (B,num_classes)."""
    sd = _strip(sd)
    x = feats.transpose(1, 2)                                   # (B,C,T), model.py:35
    for i in range(layers):
        d = 2 ** i
        p = 'tcn.mb_ms_tcn.network.%d.' % i

        def branches(inp, stage):
            outs = []
            for k_idx, k in enumerate(ksizes):
                q = p + 'cbcr%d_%d.' % (stage, k_idx)
                pad = (k - 1) * d
                o = F.conv1d(inp, sd[q + 'conv.weight'], sd[q + 'conv.bias'], dilation=d, padding=pad)
                o = _bn(o, sd, q + 'batchnorm')
                if pad:
                    o = o[:, :, pad // 2:-(pad // 2)]          # Chomp1d(symm_chomp=True), tcn.py:22-23
                outs.append(F.prelu(o, sd[q + 'non_lin.weight']))
            return torch.cat(outs, 1)
        out1 = branches(branches(x, 0), 1)
        res = x
        if p + 'downsample.weight' in sd:
            res = F.conv1d(x, sd[p + 'downsample.weight'], sd[p + 'downsample.bias'])
        x = F.prelu(out1 + res, sd[p + 'relu_final.weight'])
    pooled = torch.stack([x[b, :, :int(l)].mean(dim=1) for b, l in enumerate(lengths)], 0)   # model.py:16-17
    return F.linear(pooled, sd['tcn.tcn_output.weight'], sd['tcn.tcn_output.bias'])


def temporal_mean(feats, lengths=None):
    """train_fusion.py:400 (mean over frames of one clip); batched form
    models/video_models/model.py:16-17 averages the first len_i frames."""
    if lengths is None:
        return feats.mean(dim=1)
    return torch.stack([feats[i, :int(l)].mean(dim=0) for i, l in enumerate(lengths)], 0)


# ----------------------------------------------------------------------------- fusion
def feature_normalize_torch(data):
    """train_fusion.py:233-238: row-wise z-norm, torch.std (unbiased), no eps."""
    mu = data.mean(dim=1, keepdim=True)
    sd_ = data.std(dim=1, keepdim=True)
    return (data - mu) / sd_


def concat_fusion(xv_audio, em_video):
    """train_fusion.py:405-410: z-norm each modality, cat([audio, video], 1)."""
    return torch.cat([feature_normalize_torch(xv_audio), feature_normalize_torch(em_video)], dim=1)


def linearfusion_forward(sd, x, extract_feats):
    """models/fusion_models/model_fusion.py:19-24."""
    sd = _strip(sd)
    x1 = F.leaky_relu(_bn(F.linear(x, sd['fc1.weight'], sd['fc1.bias']), sd, 'bn1'), 0.2)
    return x1 if extract_feats else F.linear(x1, sd['fc2.weight'], sd['fc2.bias'])


def lowfer_forward(e1, e2):
    """models/fusion_models/LBP.py:28-54: the bilinear MFB term is computed and
    discarded; the returned value is cat([e1, sigmoid(e2), e1*sigmoid(e2)], 1)."""
    s = torch.sigmoid(e2)
    return torch.cat([e1, s, s * e1], dim=1)
