"""Generate tests/golden/*.npz by running the REFERENCE's own modules
(imported from /root/reference -- only possible in the build container) on seeded
synthetic inputs and weights from deeplip_b200.synth.  Weights/inputs are not
stored: they regenerate bit-identically from NumPy seeds; only small outputs are.

    python oracle/gen_golden.py
"""
import os, sys, hashlib
sys.dont_write_bytecode = True          # /root/reference is read-only
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
import numpy as np
import torch

from deeplip_b200 import synth
from oracle import scoring_ref

OUT = os.path.join(ROOT, 'tests', 'golden')


def video_case():
    from models.video_models.model import Lipreading
    from models.video_models.dataloaders import get_preprocessing_pipelines
    sd = synth.make_video_state_dict(seed=1, randomize=True)
    m = Lipreading(relu_type='prelu', backbone_type='resnet', extract_feats=True,
                   tcn_options=synth.TCN_OPTIONS).eval()
    m.load_state_dict(sd, strict=False)
    crops = synth.lip_crops_u8([3, 3, 7], T=6, seed=1)                  # (3,6,96,96) u8
    pre = get_preprocessing_pipelines()['test']
    x = np.stack([pre(c).astype(np.float32) for c in crops])            # reference CPU preprocessing
    with torch.no_grad():
        y = m(torch.from_numpy(x)[:, None], lengths=[6, 6, 6])
        stem = m.frontend3D(torch.from_numpy(x)[:1, None])
    np.savez_compressed(os.path.join(OUT, 'video_small.npz'),
                        pre_sample=x[0, 0, ::8, ::8], feats=y.numpy(),
                        stem_sample=stem[0, ::8, 0, ::3, ::3].numpy())


def video_tcn_case():
    """Lipreading(extract_feats=False): trunk + MS-TCN head -> logits."""
    from models.video_models.model import Lipreading
    from models.video_models.dataloaders import get_preprocessing_pipelines
    sd = synth.make_video_state_dict(seed=1, randomize=True)
    sd.update(synth.make_tcn_state_dict(num_classes=62, seed=1, randomize=True))
    m = Lipreading(relu_type='prelu', backbone_type='resnet', num_classes=62, extract_feats=False,
                   tcn_options=synth.TCN_OPTIONS).eval()
    m.load_state_dict(sd)
    crops = synth.lip_crops_u8([3, 7], T=12, seed=2)
    pre = get_preprocessing_pipelines()['test']
    x = np.stack([pre(c).astype(np.float32) for c in crops])
    with torch.no_grad():
        y = m(torch.from_numpy(x)[:, None], lengths=[12, 9])
    np.savez_compressed(os.path.join(OUT, 'video_tcn_small.npz'), logits=y.numpy())


def audio_case():
    from models.audio_models.tdnn import SpeakerEmbNet
    out = {}
    wav = synth.speech_like_audio([3, 3, 7], nsamp=16000, seed=1)
    from oracle import frontend_np
    feats = np.stack([frontend_np.extract_feature(w).T for w in wav])   # (3,24,T) -- unpinned front end
    for arch in ('etdnn', 'tdnn'):
        for pool in ('statistic', 'attentive_statistic'):
            o = synth.audio_opts(arch, pool)
            sd = synth.make_audio_state_dict(o, seed=1, randomize=True)
            net = SpeakerEmbNet(o).eval()
            net.load_state_dict(sd)
            with torch.no_grad():
                xv, x_a = net.extract_embedding(torch.from_numpy(feats))
                fw = net(torch.from_numpy(feats))
            out['%s_%s_xv' % (arch, pool)] = xv.numpy()
            out['%s_%s_xa' % (arch, pool)] = x_a.numpy()
            out['%s_%s_fw' % (arch, pool)] = fw.numpy()
    np.savez_compressed(os.path.join(OUT, 'audio_small.npz'), **out)


def fusion_case():
    from models.fusion_models.model_fusion import model_fusion
    sd = synth.make_fusion_state_dict(seed=1)
    x = torch.from_numpy(synth.structured_embeddings([1, 1, 2, 3, 4], dim=1024, seed=1))
    out = {}
    for ef in (True, False):
        m = model_fusion(1024, 512, 62, ef).eval()
        m.load_state_dict(sd)
        with torch.no_grad():
            out['linear_%d' % ef] = m(x).numpy()
    # train_fusion.py:233-238 restated verbatim on the same tensors (Trainer is not importable, D4)
    a, v = x[:, :512], x[:, 512:]
    def fn(data):
        mu = torch.mean(data, axis=1); std = torch.std(data, axis=1)
        return ((data.transpose(0, 1) - mu) / std).transpose(0, 1)
    out['concat'] = torch.cat([fn(a), fn(v)], dim=1).numpy()
    np.savez_compressed(os.path.join(OUT, 'fusion_small.npz'), **out)


def scoring_case():
    out = {}
    for name in ('grid', 'lomgrid'):
        path = '/root/reference/database/trial_%s_v1.txt' % name
        out[name + '_sha256'] = hashlib.sha256(open(path, 'rb').read()).hexdigest()
        labels, pairs = scoring_ref.parse_trials(path)
        table, enrol, test = scoring_ref.utterance_table(pairs)
        out[name + '_n_utts'] = len(table)
        out[name + '_labels_sum'] = int(labels.sum())
        out[name + '_enrol_head'] = enrol[:64]
        out[name + '_test_head'] = test[:64]
        out[name + '_enrol_crc'] = int(np.bitwise_xor.reduce(enrol.astype(np.int64) * (np.arange(len(enrol)) + 1)))
        out[name + '_test_crc'] = int(np.bitwise_xor.reduce(test.astype(np.int64) * (np.arange(len(test)) + 1)))
        spk = [synth.speaker_of_utt(u) for u in table]
        emb = synth.structured_embeddings(spk, dim=64, seed=1, within=2.5)
        # the reference loop itself (sklearn per trial) on the first 2000 trials + EER on all
        loop = np.concatenate(scoring_ref.cosine_scores_loop(emb, enrol[:2000], test[:2000]))
        out[name + '_scores_head'] = loop
        full = scoring_ref.cosine_scores_vec(emb, enrol, test)
        eer, thr = scoring_ref.eer_from_scores(labels, list(full.astype(np.float32).reshape(-1, 1)))
        out[name + '_eer'] = eer
        out[name + '_thr'] = float(thr)
    np.savez_compressed(os.path.join(OUT, 'scoring.npz'), **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(1)
    video_case(); video_tcn_case(); audio_case(); fusion_case(); scoring_case()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
