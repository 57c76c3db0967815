import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, packing, _lib
from lin_bench import timeit
DEV = 'cuda'

def run(name, N, H, W, C, Cout):
    w = packing.pack_conv_weight(torch.randn(Cout, C, 3, 3, device=DEV) * 0.05, Cout)
    sc = torch.ones(Cout, device=DEV); sh = torch.zeros(Cout, device=DEV); sl = torch.full((Cout,), 0.2, device=DEV)
    x = torch.randn(N, H, W, C, device=DEV).to(torch.bfloat16)
    res = torch.randn(N, H, W, Cout, device=DEV).to(torch.bfloat16)
    for dbg, what in ((0, 'full'), (4, 'no epilogue'), (12, 'no epi + 1/4 MMAs')):
        _lib.set_option('dbg', dbg)
        t = timeit(lambda: ops.conv_igemm(x, w, C, Cout, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=res))
        print('%-16s %-20s %7.1f us' % (name, what, t), flush=True)
    _lib.set_option('dbg', 0)

B = 64 * 75
run('layer2 128', B, 11, 11, 128, 128)
run('layer3 256', B, 6, 6, 256, 256)
run('layer4 512', B, 3, 3, 512, 512)
