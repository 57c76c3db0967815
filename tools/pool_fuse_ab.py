"""K4 fused into layer4's last conv (dl_conv_desc.avgpool) against conv + pooling kernel: interleaved A/B of the video
branch's tail at B = 64 (conv 3x3 512->512 + residual over 4 800 frames of 3x3, then the utterance mean), L2 flushed
before every repetition.   python tools/pool_fuse_ab.py"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, _lib
N, C, B, T = 4800, 512, 64, 75
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(N, 3, 3, C, device='cuda', generator=g).to(torch.bfloat16)
res = torch.randn(N, 3, 3, C, device='cuda', generator=g).to(torch.bfloat16)
w = (torch.randn(C, 9 * C, device='cuda', generator=g) * 0.02).to(torch.bfloat16)
sc = torch.rand(C, device='cuda') + 0.5
sh = torch.randn(C, device='cuda') * 0.1
sl = torch.full((C,), 0.25, device='cuda')
flush = torch.empty(160 << 20, dtype=torch.uint8, device='cuda')
def tail(fuse):
    _lib.set_option('pool_fuse', fuse)
    _, p = ops.conv_igemm(x, w, C, C, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=res, avgpool=True)
    return ops.temporal_mean(p, B, T)
def old():
    y, _ = ops.conv_igemm(x, w, C, C, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=res)
    return ops.frame_pool_temporal_mean(y, B, T, want_frames=False, want_mean=True)[1]
fns = {'fused (conv with K4 in the epilogue + temporal mean)': lambda: tail(1),
       'conv + per-image pooling kernel + temporal mean': lambda: tail(0),
       'conv + one-kernel K4 (before this change)': old}
for f in fns.values():
    for _ in range(3):
        f()
torch.cuda.synchronize()
ev = {k: [] for k in fns}
for i in range(15):
    for k, f in fns.items():
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record()
        ev[k].append((a, b))
torch.cuda.synchronize()
_lib.set_option('pool_fuse', 1)
for k, v in ev.items():
    print('%-56s median %6.1f us' % (k, statistics.median(a.elapsed_time(b) for a, b in v) * 1e3))
print('bit-identical:', torch.equal(tail(1), old()))
