"""Interleaved A/B of programmatic dependent launch (dl_set_option("pdl")) on the whole AV extraction step, B=64."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from deeplip_b200 import _lib
from deeplip_b200.pipeline import AVExtractor, build_models
dev = torch.device('cuda')
flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
audio, video = build_models(dev, seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(64, seed=1)
raw, wav = torch.from_numpy(raw).to(dev), torch.from_numpy(wav).to(dev)
outs = {}
for v in (0, 1):
    _lib.set_option('pdl', v)
    for _ in range(3):
        outs[v] = ex.extract(wav, raw).clone()
torch.cuda.synchronize()
print('bitwise equal:', torch.equal(outs[0], outs[1]))
res = {0: [], 1: []}
for rnd in range(8):
    for v in (0, 1):
        _lib.set_option('pdl', v)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            flush.zero_()
            ex.extract(wav, raw)
        b.record()
        torch.cuda.synchronize()
        res[v].append(a.elapsed_time(b) / 5)
for v in (0, 1):
    print('pdl=%d: median %.1f us/step  rounds %s' % (v, statistics.median(res[v]) * 1e3, ' '.join('%.0f' % (x * 1e3) for x in res[v])))
_lib.set_option('pdl', 1)
