"""A/B: TMA im2col operand A vs guarded-linear (tiled TMA) operand A on the trunk's stride-1 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, packing

DEV = 'cuda'


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def run(name, N, H, W, C, Cout, R, S, pad, dil=(1, 1), guard=(1, 1)):
    w = packing.pack_conv_weight(torch.randn(Cout, C, R, S, device=DEV) * 0.05, Cout)
    sc = torch.ones(Cout, device=DEV); sh = torch.zeros(Cout, device=DEV); sl = torch.full((Cout,), 0.2, device=DEV)
    x = torch.randn(N, H, W, C, device=DEV).to(torch.bfloat16)
    P = H + 2 * pad[0] - dil[0] * (R - 1); Q = W + 2 * pad[1] - dil[1] * (S - 1)
    res = torch.randn(N, P, Q, Cout, device=DEV).to(torch.bfloat16)
    t_im = timeit(lambda: ops.conv_igemm(x, w, C, Cout, R, S, (1, 1), pad, dil, sc, sh, sl, residual=res))
    Hg, Wg = H + guard[0], W + guard[1]
    xg = torch.zeros(N, Hg, Wg, C, device=DEV, dtype=torch.bfloat16); xg[:, :H, :W] = x
    rg = torch.zeros(N, Hg, Wg, Cout, device=DEV, dtype=torch.bfloat16)
    og = torch.zeros_like(rg)
    t_lin = timeit(lambda: ops.conv_igemm_lin(xg, w, C, Cout, (P, Q), R, S, pad, dil, sc, sh, sl, residual=rg, out=og))
    fl = 2.0 * N * P * Q * Cout * C * R * S
    print('%-22s im2col %7.1f us (%5.0f TF)   linear %7.1f us (%5.0f TF useful)' %
          (name, t_im, fl / t_im / 1e6, t_lin, fl / t_lin / 1e6), flush=True)


if __name__ == '__main__' and len(sys.argv) == 1:
    B = 64 * 75
    run('layer2 3x3 128', B, 11, 11, 128, 128, 3, 3, (1, 1))
    run('layer3 3x3 256', B, 6, 6, 256, 256, 3, 3, (1, 1))
    run('layer1 3x3 64', B, 22, 22, 64, 64, 3, 3, (1, 1))
    run('tdnn k5 512', 64, 1, 296, 512, 512, 1, 5, (0, 0), guard=(0, 0))
    run('tdnn k3d2 512', 64, 1, 292, 512, 512, 1, 3, (0, 0), dil=(1, 2), guard=(0, 0))
    run('tdnn k1 512', 64, 1, 288, 512, 512, 1, 1, (0, 0), guard=(0, 0))
    run('tdnn k1 512->1504', 64, 1, 282, 512, 1504, 1, 1, (0, 0), guard=(0, 0))


def ablate_lin(N=64 * 75, H=11, W=11, C=128, Cout=128):
    """Guarded-linear tap-sharing kernel with parts switched off (dl_set_option('dbg', ...))."""
    from deeplip_b200 import _lib
    w = packing.pack_conv_weight(torch.randn(Cout, C, 3, 3, device=DEV) * 0.05, Cout)
    sc = torch.ones(Cout, device=DEV); sh = torch.zeros(Cout, device=DEV); sl = torch.full((Cout,), 0.2, device=DEV)
    xg = torch.zeros(N, H + 1, W + 1, C, device=DEV, dtype=torch.bfloat16)
    xg[:, :H, :W] = torch.randn(N, H, W, C, device=DEV).to(torch.bfloat16)
    rg = torch.zeros(N, H + 1, W + 1, Cout, device=DEV, dtype=torch.bfloat16)
    og = torch.zeros_like(rg)
    for dbg, what in ((0, 'full'), (1, 'no residual'), (2, 'no stores'), (4, 'no epilogue'), (12, 'no epilogue, 1/4 MMAs')):
        _lib.set_option('dbg', dbg)
        t = timeit(lambda: ops.conv_igemm_lin(xg, w, C, Cout, (H, W), 3, 3, (1, 1), (1, 1), sc, sh, sl, residual=rg, out=og))
        print('layer2 guarded-linear taps=3  %-24s %7.1f us' % (what, t), flush=True)
    _lib.set_option('dbg', 0)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'ablate':
    ablate_lin()
