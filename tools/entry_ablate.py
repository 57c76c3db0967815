"""Ablation of the trunk convs that are not tensor-bound: the fused layer2 entry (64 -> 256, 3x3 s2, one resident
256-wide tile on the CTA-pair kernel) and a layer2 conv (128 -> 128 + residual), with parts of the kernel switched off
through dl_set_option('dbg', ...): 1 = no residual read, 2 = no stores, 4 = no epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import _lib, ops, packing
from lin_bench import timeit

DEV = 'cuda'
N = 64 * 75


def entry():
    w1 = torch.randn(128, 64, 3, 3, device=DEV) * 0.05
    wd = torch.randn(128, 64, 1, 1, device=DEV) * 0.05
    wf = torch.zeros((256, 9 * 64), device=DEV, dtype=torch.bfloat16)
    wf[:128] = packing.pack_conv_weight(w1)
    wf[128:, 4 * 64:5 * 64] = packing.pack_conv_weight(wd)
    sc = torch.ones(256, device=DEV); sh = torch.zeros(256, device=DEV); sl = torch.full((256,), 0.2, device=DEV)
    x = torch.zeros(N, 23, 22, 64, device=DEV, dtype=torch.bfloat16)
    x[:, :22] = torch.randn(N, 22, 22, 64, device=DEV).to(torch.bfloat16)
    for hint in (0, 128):
        for dbg, what in ((0, 'full'), (2, 'no stores'), (4, 'no epilogue')):
            _lib.set_option('dbg', dbg)
            t = timeit(lambda: ops.conv_igemm(x, wf, 64, 256, 3, 3, (2, 2), (1, 1), (1, 1), sc, sh, sl, H=22, W=22,
                                              center_only_from=hint))
            print('layer2 entry 64->256 s2  center_only_from=%-3d  %-12s %7.1f us' % (hint, what, t), flush=True)
    _lib.set_option('dbg', 0)


def layer2():
    w = packing.pack_conv_weight(torch.randn(128, 128, 3, 3, device=DEV) * 0.05)
    sc = torch.ones(128, device=DEV); sh = torch.zeros(128, device=DEV); sl = torch.full((128,), 0.2, device=DEV)
    x = torch.randn(N, 11, 11, 128, device=DEV).to(torch.bfloat16)
    res = torch.randn(N, 11, 11, 128, device=DEV).to(torch.bfloat16)
    for dbg, what in ((0, 'full'), (1, 'no residual'), (2, 'no stores'), (3, 'no residual, no stores'), (4, 'no epilogue')):
        _lib.set_option('dbg', dbg)
        t = timeit(lambda: ops.conv_igemm(x, w, 128, 128, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=res))
        print('layer2 3x3 128->128 +res  %-24s %7.1f us' % (what, t), flush=True)
    _lib.set_option('dbg', 0)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        _lib.set_option('staged_epilogue', int(sys.argv[1]))
        print('-- staged_epilogue =', sys.argv[1])
    entry()
    layer2()
