import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, packing
from lin_bench import timeit
DEV = 'cuda'
N, H, W = 64 * 75, 22, 22
w = packing.pack_conv_weight(torch.randn(64, 64, 3, 3, device=DEV) * 0.05, 64)
sc = torch.ones(64, device=DEV); sh = torch.zeros(64, device=DEV); sl = torch.full((64,), 0.2, device=DEV)
x = torch.zeros(N, H + 1, W, 64, device=DEV, dtype=torch.bfloat16)
x[:, :H] = torch.randn(N, H, W, 64, device=DEV).to(torch.bfloat16)
r = torch.zeros_like(x); r[:, :H] = torch.randn(N, H, W, 64, device=DEV).to(torch.bfloat16)
out = torch.zeros_like(x)
fl = 2.0 * N * H * W * 64 * 64 * 9
for name, res in (('halo no residual', None), ('halo + residual', r)):
    t = timeit(lambda: ops.conv3x3_halo(x, w, sc, sh, sl, H, out, residual=res))
    print('%-18s %7.1f us (%5.0f TF useful)' % (name, t, fl / t / 1e6), flush=True)
