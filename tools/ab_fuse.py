"""Interleaved A/B of the fused layer2 entry block (video branch, B=64): alternating rounds to cancel clock drift."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from deeplip_b200.pipeline import build_models
from deeplip_b200.video_models import resnet as R
dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
audio, video = build_models(dev, seed=1)
raw, _ = bench.synth_batch(64, seed=1)
raw = torch.from_numpy(raw).to(dev)


def once():
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); video.utterance_embedding(raw); b.record()
    return a, b


res = {False: [], True: []}
for f in (False, True):
    R.FUSE_L2_ENTRY = f
    for _ in range(3):
        video.utterance_embedding(raw)
torch.cuda.synchronize()
for rnd in range(8):
    for f in (False, True):
        R.FUSE_L2_ENTRY = f
        evs = [once() for _ in range(5)]
        torch.cuda.synchronize()
        res[f].append(statistics.median(a.elapsed_time(b) for a, b in evs))
for f in (False, True):
    print('fused=%d: median %.1f us  rounds %s' % (f, statistics.median(res[f]) * 1e3, ' '.join('%.0f' % (v * 1e3) for v in res[f])))
