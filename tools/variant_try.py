"""A/B of kernel-selection switches (dl_set_option) on the layer2 conv shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import _lib, ops, packing
from lin_bench import timeit
DEV = 'cuda'
N = 64 * 75
w = packing.pack_conv_weight(torch.randn(128, 128, 3, 3, device=DEV) * 0.05)
sc = torch.ones(128, device=DEV); sh = torch.zeros(128, device=DEV); sl = torch.full((128,), 0.2, device=DEV)
x = torch.randn(N, 11, 11, 128, device=DEV).to(torch.bfloat16)
res = torch.randn(N, 11, 11, 128, device=DEV).to(torch.bfloat16)
for opts in ({}, {'staged_epilogue': 0}, {'pair_resident': 0}):
    extra = 256 if 'dbg256' in opts else 0
    for k, v in opts.items():
        if k != 'dbg256':
            _lib.set_option(k, v)
    for dbg, what in ((0, 'full'), (1, 'no residual'), (4, 'no epilogue')):
        _lib.set_option('dbg', dbg | extra)
        t = timeit(lambda: ops.conv_igemm(x, w, 128, 128, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=res))
        print('layer2 3x3 128->128 +res  %-24s %-12s %7.1f us' % (opts, what, t), flush=True)
    _lib.set_option('dbg', 0)
    for k in opts:
        if k != 'dbg256':
            _lib.set_option(k, 1)
