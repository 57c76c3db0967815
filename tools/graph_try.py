import sys, time; sys.path.insert(0,'/root/repo')
import torch, bench
from deeplip_b200.pipeline import AVExtractor, GraphedExtractor, build_models
from deeplip_b200 import _lib
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(64, seed=1)
raw, wav = torch.from_numpy(raw).cuda(), torch.from_numpy(wav).cuda()
ref = ex.extract(wav, raw).clone()
g = GraphedExtractor(ex, wav, raw)
out = g.extract(wav, raw)
torch.cuda.synchronize()
print('graph vs eager max abs', float((out - ref).abs().max()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, n=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): fn()
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        flush.zero_(); fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print('eager ms/step', timeit(lambda: ex.extract(wav, raw)))
print('graph ms/step', timeit(lambda: g.extract(wav, raw)))
