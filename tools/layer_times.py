"""Per-launch CUDA-event times of one extraction step (B utterances): every ops.* call of the video and audio branches
is bracketed by events on the launching stream; medians over `reps` steps after warm-up, L2 flushed between steps.
    python tools/layer_times.py [B] [reps]"""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import AVExtractor, build_models

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device('cuda', 0)
audio, video = build_models(dev, seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(B, seed=1)
raw, wav = torch.from_numpy(raw).to(dev), torch.from_numpy(wav).to(dev)
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
log = []
NAMES = ['frontend_features', 'stem_conv3d', 'conv_igemm', 'conv_igemm_lin', 'conv3x3_halo', 'frame_pool_temporal_mean',
         'stat_pool', 'znorm_concat', 'affine_act']


def wrap(name):
    fn = getattr(ops, name)

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        desc = name
        if name == 'conv_igemm':
            x = a[0]
            desc += ' %s Cin%d Cout%d %dx%d s%s' % (tuple(x.shape[:3]), a[2], a[3], a[4] if len(a) > 4 else k.get('R', 1),
                                                    a[5] if len(a) > 5 else k.get('S', 1), (a[6] if len(a) > 6 else k.get('stride', (1, 1)))[0])
            if k.get('residual') is not None or (len(a) > 12 and a[12] is not None):
                desc += ' +res'
        elif name == 'conv3x3_halo':
            desc += ' +res' if k.get('residual') is not None else ''
        log.append((desc, e0, e1))
        return out
    setattr(ops, name, w)


for _ in range(3):
    ex.extract(wav, raw)
for n in NAMES:
    wrap(n)
runs = []
for _ in range(reps):
    flush.zero_()
    log.clear()
    ex.extract(wav, raw)
    torch.cuda.synchronize()
    runs.append([(d, a.elapsed_time(b) * 1e3) for d, a, b in log])
tot = 0.0
for i, (d, _) in enumerate(runs[0]):
    us = statistics.median(r[i][1] for r in runs)
    tot += us
    print('%3d %8.1f us  %s' % (i, us, d))
print('sum %.1f us over %d launches-groups, B=%d' % (tot, len(runs[0]), B))
