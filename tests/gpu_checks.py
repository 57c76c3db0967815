"""Kernel-level parity checks: CUDA path (through the C ABI) vs the CPU oracle on seeded inputs.
Each check returns a dict of error metrics and raises AssertionError when out of tolerance.
Used by tests/test_gpu_*.py and tools/gpu_diag.py."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from deeplip_b200 import ops, packing, synth
from oracle import frontend_np, models_ref, scoring_ref

DEV = 'cuda'


def bf16r(t):
    return t.to(torch.bfloat16).float()


def rel_err(a, b):
    a = a.float().cpu()
    b = b.float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def conv_case(N, H, W, C, Cout, R, S, stride=1, pad=0, dil=(1, 1), residual=False, f32=False, seed=0, ld=None,
              tol=1.5e-2, small_linear=1):
    """small_linear: dl_set_option("small_linear") -- 1 routes fc layers (P Q == 1, up to 4096 rows) to
    linear_small_kernel, 0 keeps them on the tensor-core igemm."""
    from deeplip_b200 import _lib
    _lib.set_option('small_linear', small_linear)
    try:
        return _conv_case(N, H, W, C, Cout, R, S, stride, pad, dil, residual, f32, seed, ld, tol)
    finally:
        _lib.set_option('small_linear', 1)


def _conv_case(N, H, W, C, Cout, R, S, stride, pad, dil, residual, f32, seed, ld, tol):
    g = torch.Generator().manual_seed(seed)
    ld = ld or packing.ceil_to(C, 8)
    x = bf16r(torch.randn(N, H, W, C, generator=g))
    w = bf16r(torch.randn(Cout, C, R, S, generator=g) / (C * R * S) ** 0.5)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.2
    slope = torch.rand(Cout, generator=g) * 0.5
    acc = F.conv2d(x.permute(0, 3, 1, 2), w, None, stride=stride, padding=pad, dilation=dil)   # (N,Cout,P,Q)
    ref = acc * scale[None, :, None, None] + shift[None, :, None, None]
    res = None
    if residual:
        res = bf16r(torch.randn(ref.shape, generator=g))
        ref = ref + res
    ref = torch.where(ref > 0, ref, ref * slope[None, :, None, None])
    xpad = torch.zeros(N, H, W, ld)
    xpad[..., :C] = x
    coutp = packing.ceil_to(Cout, 8)
    st = (stride, stride) if isinstance(stride, int) else stride
    pd = (pad, pad) if isinstance(pad, int) else pad
    y, yf = ops.conv_igemm(
        xpad.to(DEV).to(torch.bfloat16), packing.pack_conv_weight(w.to(DEV), coutp), C, coutp, R, S, st, pd, dil,
        packing.pad_vec(scale.to(DEV), coutp), packing.pad_vec(shift.to(DEV), coutp),
        packing.pad_vec(slope.to(DEV), coutp),
        residual=(None if res is None else _pad_last(res.permute(0, 2, 3, 1), coutp).to(DEV).to(torch.bfloat16).contiguous()),
        want_f32=f32)
    torch.cuda.synchronize()
    out = {}
    got = y.float().cpu()[..., :Cout].permute(0, 3, 1, 2)
    out['bf16_rel'] = rel_err(got, ref)
    if coutp != Cout:
        out['pad_abs'] = float(y.float().cpu()[..., Cout:].abs().max())
    if f32:
        P, Q = acc.shape[2], acc.shape[3]
        gotf = yf.cpu().view(N, P, Q, coutp)[..., :Cout].permute(0, 3, 1, 2)
        out['f32_rel'] = rel_err(gotf, acc)
        assert out['f32_rel'] < 2e-3, out
    assert out['bf16_rel'] < tol, out
    assert out.get('pad_abs', 0.0) == 0.0, out
    return out


def conv_lin_case(N, H, W, C, Cout, R, S, pad=(0, 0), dil=(1, 1), guard=(0, 0), residual=False, seed=0, tol=1.5e-2):
    """Guarded-linear operand A (tiled TMA) vs F.conv2d: stride 1, guards of `guard` rows / columns."""
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, H, W, C, generator=g))
    w = bf16r(torch.randn(Cout, C, R, S, generator=g) / (C * R * S) ** 0.5)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.2
    slope = torch.rand(Cout, generator=g) * 0.5
    acc = F.conv2d(x.permute(0, 3, 1, 2), w, None, stride=1, padding=pad, dilation=dil)   # (N,Cout,P,Q)
    P, Q = acc.shape[2], acc.shape[3]
    ref = acc * scale[None, :, None, None] + shift[None, :, None, None]
    Hg, Wg = H + guard[0], W + guard[1]
    res_g = None
    if residual:
        res = bf16r(torch.randn(ref.shape, generator=g))
        ref = ref + res
        res_g = torch.zeros(N, Hg, Wg, Cout)
        res_g[:, :P, :Q] = res.permute(0, 2, 3, 1)
        res_g = res_g.to(DEV).to(torch.bfloat16)
    ref = torch.where(ref > 0, ref, ref * slope[None, :, None, None])
    ld = packing.ceil_to(C, 8)
    xg = torch.zeros(N, Hg, Wg, ld)
    xg[:, :H, :W, :C] = x
    y = ops.conv_igemm_lin(xg.to(DEV).to(torch.bfloat16), packing.pack_conv_weight(w.to(DEV), Cout), C, Cout, (P, Q),
                           R, S, pad, dil, scale.to(DEV), shift.to(DEV), slope.to(DEV), residual=res_g)
    torch.cuda.synchronize()
    yc = y.float().cpu()
    out = {'bf16_rel': rel_err(yc[:, :P, :Q].permute(0, 3, 1, 2), ref)}
    mask = torch.ones(Hg, Wg, dtype=torch.bool)
    mask[:P, :Q] = False
    out['guard_abs'] = float(yc[:, mask].abs().max()) if mask.any() else 0.0
    assert out['bf16_rel'] < tol, out
    assert out['guard_abs'] == 0.0, out
    return out


def conv_guarded_io_case(N=3, H=22, W=22, C=64, Cout=128, R=3, S=3, stride=2, pad=1, seed=0, tol=1.5e-2):
    """im2col kernel reading a guarded input (img_rows / img_cols) and writing a guarded output."""
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, H, W, C, generator=g))
    w = bf16r(torch.randn(Cout, C, R, S, generator=g) / (C * R * S) ** 0.5)
    scale = torch.rand(Cout, generator=g) + 0.5
    shift = torch.randn(Cout, generator=g) * 0.2
    slope = torch.rand(Cout, generator=g) * 0.5
    acc = F.conv2d(x.permute(0, 3, 1, 2), w, None, stride=stride, padding=pad)
    P, Q = acc.shape[2], acc.shape[3]
    ref = acc * scale[None, :, None, None] + shift[None, :, None, None]
    ref = torch.where(ref > 0, ref, ref * slope[None, :, None, None])
    xg = torch.zeros(N, H + 1, W + 2, C)
    xg[:, :H, :W] = x
    out_buf = torch.zeros(N, P + 1, Q + 1, Cout, device=DEV, dtype=torch.bfloat16)
    y, _ = ops.conv_igemm(xg.to(DEV).to(torch.bfloat16), packing.pack_conv_weight(w.to(DEV), Cout), C, Cout, R, S,
                          (stride, stride), (pad, pad), (1, 1), scale.to(DEV), shift.to(DEV), slope.to(DEV),
                          H=H, W=W, out=out_buf)
    torch.cuda.synchronize()
    yc = y.float().cpu()
    out = {'bf16_rel': rel_err(yc[:, :P, :Q].permute(0, 3, 1, 2), ref),
           'guard_abs': float(max(yc[:, P:].abs().max(), yc[:, :, Q:].abs().max()))}
    assert out['bf16_rel'] < tol and out['guard_abs'] == 0.0, out
    return out


def halo_case(N=5, H=22, W=22, residual=True, seed=0):
    """dl_conv3x3_c64_halo_bf16 on the stacked-rows layout vs F.conv2d."""
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, H, W, 64, generator=g))
    w = bf16r(torch.randn(64, 64, 3, 3, generator=g) / 24.0)
    scale = torch.rand(64, generator=g) + 0.5
    shift = torch.randn(64, generator=g) * 0.2
    slope = torch.rand(64, generator=g) * 0.5
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    res = bf16r(torch.randn(N, H, W, 64, generator=g)) if residual else None
    if residual:
        ref = ref + res.permute(0, 3, 1, 2)
    ref = torch.where(ref > 0, ref, ref * slope[None, :, None, None]).permute(0, 2, 3, 1)

    def stack(t):
        o = torch.zeros(N, H + 1, W, 64)
        o[:, :H] = t
        return o.to(DEV).to(torch.bfloat16)
    out = torch.zeros(N, H + 1, W, 64, device=DEV, dtype=torch.bfloat16)
    ops.conv3x3_halo(stack(x), packing.pack_conv_weight(w.to(DEV)), scale.to(DEV), shift.to(DEV), slope.to(DEV), H,
                     out=out, residual=None if res is None else stack(res))
    torch.cuda.synchronize()
    got = out.float().cpu()
    m = {'rel': rel_err(got[:, :H], ref), 'pad_abs': float(got[:, H:].abs().max())}
    d = (got[:, :H] - ref).abs()
    m['worst'] = [int(v) for v in np.unravel_index(int(d.argmax()), d.shape)]
    assert m['rel'] < 1.5e-2 and m['pad_abs'] == 0.0, m
    return m


def _pad_last(t, n):
    if t.shape[-1] == n:
        return t
    out = torch.zeros(*t.shape[:-1], n)
    out[..., :t.shape[-1]] = t
    return out


CONV_CASES = {
    'l1_3x3_64': dict(N=3, H=22, W=22, C=64, Cout=64, R=3, S=3, stride=1, pad=1, residual=True),
    'l2_3x3s2_64_128': dict(N=3, H=22, W=22, C=64, Cout=128, R=3, S=3, stride=2, pad=1),
    'l2_1x1s2_64_128': dict(N=3, H=22, W=22, C=64, Cout=128, R=1, S=1, stride=2, pad=0),
    'l2_3x3_128': dict(N=2, H=11, W=11, C=128, Cout=128, R=3, S=3, stride=1, pad=1, residual=True),
    'l3_3x3s2_128_256': dict(N=2, H=11, W=11, C=128, Cout=256, R=3, S=3, stride=2, pad=1),
    'l3_3x3_256': dict(N=5, H=6, W=6, C=256, Cout=256, R=3, S=3, stride=1, pad=1, residual=True),
    'l4_3x3s2_256_512': dict(N=5, H=6, W=6, C=256, Cout=512, R=3, S=3, stride=2, pad=1),
    'l4_3x3_512': dict(N=40, H=3, W=3, C=512, Cout=512, R=3, S=3, stride=1, pad=1, residual=True),
    'tdnn_k5_24_512': dict(N=2, H=1, W=50, C=64, Cout=512, R=1, S=5),
    'tdnn_k3d3_512': dict(N=2, H=1, W=140, C=512, Cout=512, R=1, S=3, dil=(1, 3)),
    'tdnn_k1_512_1500': dict(N=2, H=1, W=130, C=512, Cout=1500, R=1, S=1),
    'fc_3000_512': dict(N=5, H=1, W=1, C=3000, Cout=512, R=1, S=1, f32=True),
    'fc_3000_512_igemm': dict(N=5, H=1, W=1, C=3000, Cout=512, R=1, S=1, f32=True, small_linear=0),
    'fc_3000_512_b64': dict(N=64, H=1, W=1, C=3000, Cout=512, R=1, S=1, f32=True),
    'fc_512_512_b100': dict(N=100, H=1, W=1, C=512, Cout=512, R=1, S=1, f32=True),
    'fc_1024_512_b33_ld': dict(N=33, H=1, W=1, C=1024, Cout=504, R=1, S=1, f32=True, ld=1088),
    'fc_512_512_b300': dict(N=300, H=1, W=1, C=512, Cout=512, R=1, S=1, f32=True),
    'conv1x1_small_map': dict(N=2, H=3, W=5, C=64, Cout=128, R=1, S=1),
    'many_tiles': dict(N=64, H=22, W=22, C=64, Cout=64, R=3, S=3, stride=1, pad=1, residual=True),
    'pair_128': dict(N=300, H=11, W=11, C=128, Cout=128, R=3, S=3, stride=1, pad=1, residual=True),
    'pair_256_s2': dict(N=1101, H=11, W=11, C=128, Cout=256, R=3, S=3, stride=2, pad=1),
    'pair_512_odd': dict(N=3700, H=3, W=3, C=256, Cout=512, R=3, S=3, stride=1, pad=1, residual=True),
    # staged (TMA store) epilogue of the streaming pair kernel with a last slab that is half outside Cout (1504 = 23.5 x 64)
    'pair_tdnn_k1_1504': dict(N=70, H=1, W=280, C=512, Cout=1500, R=1, S=1),
    'pair_tdnn_k3_res_odd_rows': dict(N=67, H=1, W=293, C=512, Cout=512, R=1, S=3, dil=(1, 2), residual=True),
}


def conv_center_only_case(N=150, H=22, W=22, seed=5):
    """dl_conv_desc.center_only_from: a 64 -> 256 3x3 stride-2 conv whose channels >= 128 carry a 1x1 conv on the centre
    tap (layer2's fused entry block) at a size that runs on the CTA-pair kernel: with the hint (two N = 128 MMA streams,
    compact weights, staged epilogue) == without it (N = 256 MMAs over the zero weights, per-lane stores with
    staged_epilogue = 0) bit for bit, and both == the fp32 reference within tolerance."""
    from deeplip_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(N, H, W, 64, generator=g))
    w1 = bf16r(torch.randn(128, 64, 3, 3, generator=g) / 24.0)
    wd = bf16r(torch.randn(128, 64, 1, 1, generator=g) / 8.0)
    scale, shift, slope = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g) * 0.2, torch.rand(256, generator=g) * 0.5
    ref = torch.cat([F.conv2d(x.permute(0, 3, 1, 2), w1, None, stride=2, padding=1),
                     F.conv2d(x.permute(0, 3, 1, 2), wd, None, stride=2, padding=0)], dim=1)
    ref = ref * scale[None, :, None, None] + shift[None, :, None, None]
    ref = torch.where(ref > 0, ref, ref * slope[None, :, None, None])
    wf = torch.zeros((256, 9 * 64), device=DEV, dtype=torch.bfloat16)
    wf[:128] = packing.pack_conv_weight(w1.to(DEV))
    wf[128:, 4 * 64:5 * 64] = packing.pack_conv_weight(wd.to(DEV))
    xd = x.to(DEV).to(torch.bfloat16)
    args = (xd, wf, 64, 256, 3, 3, (2, 2), (1, 1), (1, 1), scale.to(DEV), shift.to(DEV), slope.to(DEV))
    hinted, _ = ops.conv_igemm(*args, center_only_from=128)
    _lib.set_option('staged_epilogue', 0)
    try:
        plain, _ = ops.conv_igemm(*args)
    finally:
        _lib.set_option('staged_epilogue', 1)
    torch.cuda.synchronize()
    out = {'equal': bool(torch.equal(hinted, plain)), 'rel': rel_err(hinted.float().cpu().permute(0, 3, 1, 2), ref)}
    assert out['equal'] and out['rel'] < 1.5e-2, out
    return out


def staged_epilogue_bitwise_case(seed=9):
    """The staged (shared memory + TMA store) epilogue against the per-lane stores, bit for bit, on the three kernel
    variants that have it: resident 128-wide (+ residual, output in a wider-pitch buffer), streaming 256-wide with two n
    blocks (+ residual), streaming with a split output."""
    from deeplip_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    out = {}

    def both(fn):
        a = fn()
        _lib.set_option('staged_epilogue', 0)
        try:
            b = fn()
        finally:
            _lib.set_option('staged_epilogue', 1)
        torch.cuda.synchronize()
        return a, b

    def prm(c):
        return (torch.rand(c, generator=g) + 0.5).to(DEV), (torch.randn(c, generator=g) * 0.2).to(DEV), (torch.rand(c, generator=g) * 0.5).to(DEV)
    # resident 128-wide tile writing channels [0,128) of a pitch-256 buffer, residual read from channels [128,256)
    N = 310
    x = torch.randn(N, 11, 11, 256, generator=g).to(DEV).to(torch.bfloat16)
    w = packing.pack_conv_weight((torch.randn(128, 128, 3, 3, generator=g) / 34.0).to(DEV))
    sc, sh, sl = prm(128)

    def l2():
        o = torch.zeros(N, 11, 11, 256, device=DEV, dtype=torch.bfloat16)
        ops.conv_igemm(x, w, 128, 128, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=x, residual_channel_offset=128, out=o)
        return o
    a, b = both(l2)
    out['resident_128_pitched'] = bool(torch.equal(a, b)) and float(a[..., 128:].abs().max()) == 0.0
    # streaming 256-wide, two n blocks, residual, M not a multiple of 256
    x2 = torch.randn(3701, 3, 3, 256, generator=g).to(DEV).to(torch.bfloat16)
    w2 = packing.pack_conv_weight((torch.randn(512, 256, 3, 3, generator=g) / 48.0).to(DEV))
    r2 = torch.randn(3701, 3, 3, 512, generator=g).to(DEV).to(torch.bfloat16)
    sc2, sh2, sl2 = prm(512)
    a, b = both(lambda: ops.conv_igemm(x2, w2, 256, 512, 3, 3, (1, 1), (1, 1), (1, 1), sc2, sh2, sl2, residual=r2)[0])
    out['streaming_256_res'] = bool(torch.equal(a, b))
    # split output (two dense tensors, second tensor map), centre-tap-only upper half
    x3 = torch.randn(1200, 11, 11, 128, generator=g).to(DEV).to(torch.bfloat16)
    w3 = torch.zeros((512, 9 * 128), device=DEV, dtype=torch.bfloat16)
    w3[:256] = packing.pack_conv_weight((torch.randn(256, 128, 3, 3, generator=g) / 34.0).to(DEV))
    w3[256:, 4 * 128:5 * 128] = packing.pack_conv_weight((torch.randn(256, 128, 1, 1, generator=g) / 11.0).to(DEV))
    sc3, sh3, sl3 = prm(512)
    a, b = both(lambda: ops.conv_igemm(x3, w3, 128, 512, 3, 3, (2, 2), (1, 1), (1, 1), sc3, sh3, sl3, split=(256, True))[0])
    out['split_output'] = bool(torch.equal(a[0], b[0])) and bool(torch.equal(a[1], b[1]))
    assert all(out.values()), out
    return out



def stem_case(B=2, T=6, H=88, W=88, u8=False, seed=0, signs=None):
    """signs: None = positive BN scales and PReLU slopes (the second-generation stem pools the raw accumulators first, max
    only); 'mixed_scale' = negative BN weights on a few channels of the first 32 (those warps pool max AND min, the other
    warps max only); 'neg_slope' = negative PReLU slopes on a few channels of the last 32 (those warps take the general
    order: BN + PReLU on every conv value, then the max-pool)."""
    sd = synth.make_video_state_dict(seed=seed + 1, randomize=True)
    if signs == 'mixed_scale':
        sd['frontend3D.1.weight'][[1, 7, 30]] *= -1.0
    elif signs == 'neg_slope':
        sd['frontend3D.2.weight'][[33, 40, 63]] *= -1.0
        sd['frontend3D.1.weight'][[2, 35]] *= -1.0
    elif signs is not None:
        raise ValueError(signs)
    if u8:
        raw = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=H + 8, W=W + 8, seed=seed + 1))
        x = torch.stack([models_ref.video_preprocess(r, crop=H) for r in raw])        # (B,T,H,W) f32
        xin = raw.to(DEV)
    else:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, T, H, W, generator=g)
        xin = x.to(DEV)
    sdr = dict(sd)
    sdr['frontend3D.0.weight'] = bf16r(sd['frontend3D.0.weight'])
    ref = models_ref.video_frontend3d(sdr, bf16r(x)[:, None])                          # (B,64,T,H/4,W/4)
    ref = ref.permute(0, 2, 3, 4, 1).reshape(B * T, H // 4, W // 4, 64)
    w = packing.pack_stem_weight(sd['frontend3D.0.weight'].to(DEV))
    s, h = packing.fold_bn(sd['frontend3D.1.weight'], sd['frontend3D.1.bias'], sd['frontend3D.1.running_mean'],
                           sd['frontend3D.1.running_var'])
    y = ops.stem_conv3d(xin, w, s.to(DEV), h.to(DEV), sd['frontend3D.2.weight'].to(DEV), crop=(H, W))
    torch.cuda.synchronize()
    # the two pre-pass generations (dl_set_option("prepass")) must agree bit for bit
    from deeplip_b200 import _lib
    _lib.set_option('prepass', 1)
    try:
        y1 = ops.stem_conv3d(xin, w, s.to(DEV), h.to(DEV), sd['frontend3D.2.weight'].to(DEV), crop=(H, W))
        torch.cuda.synchronize()
    finally:
        _lib.set_option('prepass', 2)
    assert torch.equal(y.view(torch.int16), y1.view(torch.int16)), 'pre-pass generations differ'
    out = {'rel': rel_err(y, ref)}
    d = (y.float().cpu() - ref).abs()
    out['worst'] = [int(v) for v in np.unravel_index(int(d.argmax()), d.shape)]
    assert out['rel'] < 1.5e-2, out
    # Both kernel generations (dl_set_option("stem"): 2 = channels-on-lanes kernel where the shape allows it (W = 88,
    # H % 8 == 0), 1 = first generation for every shape) against the oracle, and against each other: they sum the same
    # products in a different order, so they agree to fp32 accumulation error, i.e. a bf16 ulp on a few values.
    _lib.set_option('stem', 1)
    try:
        yg1 = ops.stem_conv3d(xin, w, s.to(DEV), h.to(DEV), sd['frontend3D.2.weight'].to(DEV), crop=(H, W))
        torch.cuda.synchronize()
    finally:
        _lib.set_option('stem', 2)
    out['rel_gen1'] = rel_err(yg1, ref)
    assert out['rel_gen1'] < 1.5e-2, out
    dd = (y.float() - yg1.float()).abs()
    out['gen1_vs_gen2_max_abs'] = float(dd.max())
    out['gen1_vs_gen2_differing'] = int((y.view(torch.int16) != yg1.view(torch.int16)).sum())
    tol = 2.0 ** -7 * max(1.0, float(y.float().abs().max()))          # one bf16 ulp at the largest magnitude
    assert out['gen1_vs_gen2_max_abs'] <= tol and out['gen1_vs_gen2_differing'] <= max(8, y.numel() // 2000), out
    if not (W == 88 and H % 8 == 0):
        assert out['gen1_vs_gen2_differing'] == 0, out      # same kernel either way
    return out


def stem_ragged_stacked_case(B=3, T=9, seed=5):
    """Second-generation stem with ragged clip lengths into the stacked-rows layout (one zero row between frames):
    frames past a clip's length equal the result for zero normalised frames, the pad row stays zero, and the kernel
    agrees with the first generation.  Odd T: the last frame pair of a clip is a half pair."""
    sd = synth.make_video_state_dict(seed=seed, randomize=True)
    w = packing.pack_stem_weight(sd['frontend3D.0.weight'].to(DEV))
    s, h = packing.fold_bn(sd['frontend3D.1.weight'], sd['frontend3D.1.bias'], sd['frontend3D.1.running_mean'],
                           sd['frontend3D.1.running_var'])
    a = sd['frontend3D.2.weight'].to(DEV)
    raw = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=96, W=96, seed=seed)).to(DEV)
    lens = [T, max(1, T // 2), T - 2][:B]
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    from deeplip_b200 import _lib
    outs = []
    for opt in (1, 2):
        _lib.set_option('stem', opt)
        try:
            o = torch.zeros(B * T, 23, 22, 64, device=DEV, dtype=torch.bfloat16)
            ops.stem_conv3d(raw, w, s.to(DEV), h.to(DEV), a, crop=(88, 88), out=o, lengths=ln)
            torch.cuda.synchronize()
        finally:
            _lib.set_option('stem', 2)
        outs.append(o)
    out = {'max_abs_gen1_vs_gen2': float((outs[0].float() - outs[1].float()).abs().max()),
           'pad_row_zero': bool((outs[1][:, 22] == 0).all())}
    # each clip alone, truncated to its length, must give the same valid frames bit for bit (gen2)
    same = True
    for b, L in enumerate(lens):
        ob = ops.stem_conv3d(raw[b:b + 1, :L].contiguous(), w, s.to(DEV), h.to(DEV), a, crop=(88, 88))
        torch.cuda.synchronize()
        # the last two valid frames see zero frames beyond the clip in both runs (temporal padding)
        same = same and bool(torch.equal(ob.view(torch.int16), outs[1][b * T:b * T + L, :22].view(torch.int16)))
    out['ragged_equals_alone'] = same
    assert out['pad_row_zero'] and out['ragged_equals_alone'] and out['max_abs_gen1_vs_gen2'] < 0.05, out
    return out


def stat_pool_case(B=3, T=277, C=1500, seed=0, lengths=None):
    g = torch.Generator().manual_seed(seed)
    ld = packing.ceil_to(C, 8)
    x = bf16r(torch.randn(B, T, C, generator=g) * 2 + 0.5)
    xp = _pad_last(x, ld)
    ln = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=DEV)
    f32, b16 = ops.stat_pool(xp.to(DEV).to(torch.bfloat16), C, lengths=ln)
    torch.cuda.synchronize()
    if lengths is None:
        ref = models_ref.mean_std_pooling(x.permute(0, 2, 1))
    else:
        ref = torch.cat([models_ref.mean_std_pooling(x[i:i + 1, :l].permute(0, 2, 1)) for i, l in enumerate(lengths)])
    out = {'abs': float((f32.cpu() - ref).abs().max()), 'bf16_rel': rel_err(b16, ref)}
    assert out['abs'] < 2e-5 and out['bf16_rel'] < 5e-3, out
    return out


def attn_pool_case(B=3, T=120, C=512, Hd=64, seed=0):
    from deeplip_b200.audio_models.pooling import AttentiveStatPooling
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(B, C, T, generator=g))
    m = AttentiveStatPooling(C, Hd)
    sd = {'pooling.' + k: v.detach().clone() for k, v in m.state_dict().items()}
    sd['pooling.W'] = bf16r(sd['pooling.W'])
    m.load_state_dict({k[8:]: v for k, v in sd.items()})
    ref = models_ref.attentive_stat_pooling(sd, x)
    got = m.to(DEV)(x.to(DEV))
    torch.cuda.synchronize()
    out = {'abs': float((got.cpu() - ref).abs().max())}
    assert out['abs'] < 2e-3, out
    return out


def frame_pool_case(B=3, T=7, HW=9, C=512, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn(B * T, HW, C, generator=g))
    lengths = [T, max(1, T - 2), max(1, T // 2)][:B]
    ff, um = ops.frame_pool_temporal_mean(x.view(B * T, HW, 1, C).to(DEV).to(torch.bfloat16), B, T,
                                          lengths=torch.tensor(lengths, dtype=torch.int32, device=DEV))
    torch.cuda.synchronize()
    ref_f = x.mean(dim=1).view(B, T, C)
    ref_m = models_ref.temporal_mean(ref_f, lengths)
    out = {'frames_abs': float((ff.cpu() - ref_f).abs().max()), 'mean_abs': float((um.cpu() - ref_m).abs().max())}
    assert out['frames_abs'] < 1e-5 and out['mean_abs'] < 1e-5, out
    return out


def fusion_case(B=9, D=512, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, D, generator=g) * 3 + 1
    v = torch.randn(B, D, generator=g) * 0.5 - 2
    ref = models_ref.concat_fusion(a, v)
    got = ops.znorm_concat(a.to(DEV), v.to(DEV))
    ref_np = np.stack([scoring_ref.featurefusion_embedding(a[i].numpy(), v[i].numpy()) for i in range(B)])
    got_np = ops.znorm_concat(a.to(DEV), v.to(DEV), biased=True, video_first=True)
    got_l2 = ops.znorm_concat(a.to(DEV), v.to(DEV), l2norm=True)
    lf = ops.lowfer(a.to(DEV), v.to(DEV))
    l2 = ops.l2_normalize(a.to(DEV))
    torch.cuda.synchronize()
    out = {'concat_abs': float((got.cpu() - ref).abs().max()),
           'np_abs': float(np.abs(got_np.cpu().numpy() - ref_np).max()),
           'l2_abs': float((got_l2.cpu() - F.normalize(ref, dim=1)).abs().max()),
           'lowfer_abs': float((lf.cpu() - models_ref.lowfer_forward(a, v)).abs().max()),
           'l2n_abs': float((l2.cpu() - F.normalize(a, dim=1)).abs().max())}
    assert max(out.values()) < 2e-5, out
    return out


def scoring_case(n_utt=500, D=1024, n_trials=3000, seed=0):
    rng = np.random.default_rng(seed)
    emb = synth.structured_embeddings(rng.integers(0, 30, n_utt), dim=D, seed=seed + 1)
    emb[7] = 0.0                                     # zero row: sklearn divides by 1
    enrol = rng.integers(0, n_utt, n_trials).astype(np.int32)
    test = rng.integers(0, n_utt, n_trials).astype(np.int32)
    k = min(3, n_trials)
    enrol[:k], test[:k] = 7, np.array([7, 8, 9])[:k]
    got = ops.cosine_score_trials(torch.from_numpy(emb).to(DEV), torch.from_numpy(enrol).to(DEV),
                                  torch.from_numpy(test).to(DEV))
    torch.cuda.synchronize()
    ref = scoring_ref.cosine_scores_vec(emb, enrol, test)
    loop = np.concatenate(scoring_ref.cosine_scores_loop(emb, enrol[:200], test[:200]))
    out = {'abs': float(np.abs(got.cpu().numpy() - ref).max()), 'loop_abs': float(np.abs(got.cpu().numpy()[:200] - loop).max())}
    embv = synth.structured_embeddings(rng.integers(0, 30, n_utt), dim=512, seed=seed + 2)
    gf = ops.score_fusion_trials(torch.from_numpy(emb).to(DEV), torch.from_numpy(embv).to(DEV),
                                 torch.from_numpy(enrol).to(DEV), torch.from_numpy(test).to(DEV))
    reff = 0.5 * ref + 0.5 * np.array([scoring_ref.torch_cosine_eps(embv[a].astype(np.float64), embv[b].astype(np.float64))
                                       for a, b in zip(enrol, test)])
    out['fusion_abs'] = float(np.abs(gf.cpu().numpy() - reff).max())
    assert max(out.values()) < 2e-6, out
    return out


def frontend_case(B=3, nsamp=16000, feat_type='mfcc', n_feat=24, seed=0, lengths=None, gen=2, stft_pad='reflect',
                  delta=0):
    """K1 against oracle/frontend_np.py; gen selects the kernel generation (dl_set_option("frontend"))."""
    from deeplip_b200 import _lib
    wav = synth.speech_like_audio(list(range(B)), nsamp=nsamp, seed=seed + 1)
    ln = None if lengths is None else torch.tensor(lengths, dtype=torch.int32, device=DEV)
    _lib.set_option('frontend', gen)
    _lib.set_option('stft_pad', 0 if stft_pad == 'reflect' else 1)
    try:
        f32, b16 = ops.frontend_features(torch.from_numpy(wav).to(DEV), feat_type, n_feat, lengths=ln, delta=delta)
        torch.cuda.synchronize()
    finally:
        _lib.set_option('frontend', 2)
        _lib.set_option('stft_pad', 0)
    if feat_type == 'stft':
        n_feat = 257
    opts = dict(num_cep=n_feat, num_bin=n_feat, pad_mode=stft_pad, delta=delta)
    n_feat = n_feat * (1 + (2 if delta is True else int(delta)))          # rows incl. the delta features
    assert f32.shape[1] == n_feat and b16.shape[2] % 64 == 0
    assert float(b16[:, :, n_feat:].float().abs().max()) == 0.0 if b16.shape[2] > n_feat else True
    out = {'abs': 0.0}
    for i in range(B):
        n = nsamp if lengths is None else lengths[i]
        ref = frontend_np.extract_feature(wav[i, :n].astype(np.float64), 16000, feat_type, opts).T    # (F,T_i)
        got = f32[i].cpu().numpy()
        out['abs'] = max(out['abs'], float(np.abs(got[:, :ref.shape[1]] - ref).max()))
        if ref.shape[1] < got.shape[1]:
            assert np.all(got[:, ref.shape[1]:] == 0)
        gb = b16[i].float().cpu().numpy()[:ref.shape[1], :n_feat].T
        out['bf16_abs'] = max(out.get('bf16_abs', 0.0), float(np.abs(gb - ref).max()))
    assert out['abs'] < 2e-3 and out['bf16_abs'] < 5e-2, out
    return out


# ======================================================================================== model level
def cosine_rows(a, b):
    a = a.double().cpu().reshape(a.shape[0], -1)
    b = b.double().cpu().reshape(b.shape[0], -1)
    return (a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1))


def video_model_case(B=2, T=6, seed=1, u8=True, speakers=None):
    """Lipreading drop-in vs the fp32 oracle (pure fp32 weights/inputs: the real tolerance test)."""
    from deeplip_b200.video_models.model import Lipreading
    sd = synth.make_video_state_dict(seed=seed, randomize=True)
    m = Lipreading(relu_type='prelu', backbone_type='resnet', extract_feats=True, tcn_options=synth.TCN_OPTIONS)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    raw = torch.from_numpy(synth.lip_crops_u8(speakers or list(range(B)), T=T, seed=seed))
    x = torch.stack([models_ref.video_preprocess(r) for r in raw])              # (B,T,88,88)
    with torch.no_grad():
        ref = models_ref.lipreading_features(sd, x[:, None])
        got = m(x[:, None].to(DEV), lengths=[T] * B)
        emb_u8 = m.utterance_embedding(raw.to(DEV)) if u8 else None
    torch.cuda.synchronize()
    out = {'frame_cos_min': float(cosine_rows(got.reshape(B * T, -1), ref.reshape(B * T, -1)).min()),
           'rel': rel_err(got, ref)}
    if u8:
        out['utt_cos_min'] = float(cosine_rows(emb_u8, ref.mean(dim=1)).min())
        assert out['utt_cos_min'] > 0.999, out
    assert out['frame_cos_min'] > 0.999, out
    return out


def video_fused_entry_case(B=3, T=40, seed=4):
    """Entry blocks with conv1 and the 1x1 skip as ONE conv (skip weights on the centre tap; layer2: one 256-wide tile
    into a 256-pitch buffer, layer3/4: split output + centre-tap-only n blocks in the CTA-pair kernel) must equal the
    separate convs bit for bit: small batch (single-CTA igemm, the centre-tap hint ignored), medium, and a batch large
    enough (>= 1821 frames) for layer4 to run on the pair kernel."""
    from deeplip_b200.video_models import resnet as R
    from deeplip_b200.video_models.model import Lipreading
    sd = synth.make_video_state_dict(seed=seed, randomize=True)
    m = Lipreading(relu_type='prelu', backbone_type='resnet', extract_feats=True, tcn_options=synth.TCN_OPTIONS)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    saved = R.FUSE_L2_ENTRY
    out = {}
    with torch.no_grad():
        try:
            for name, (b, t) in (('small', (1, 3)), ('large', (B, T)), ('pair', (26, 72))):
                raw = torch.from_numpy(synth.lip_crops_u8(list(range(b)), T=t, seed=seed)).to(DEV)
                R.FUSE_L2_ENTRY = True
                assert m.trunk.fused_entry_enabled()
                fused = m.trunk_maps(raw).clone()
                R.FUSE_L2_ENTRY = False
                plain = m.trunk_maps(raw).clone()
                torch.cuda.synchronize()
                out[name + '_equal'] = bool(torch.equal(fused, plain))
                out[name + '_abs'] = float((fused.float() - plain.float()).abs().max())
        finally:
            R.FUSE_L2_ENTRY = saved
    assert out['small_equal'] and out['large_equal'] and out['pair_equal'], out
    return out


def video_guarded_case(B=4, T=30, seed=2, reps=8):
    """layer2 on the guarded layout (tap-sharing CTA-pair kernel) vs the dense im2col path and the fp32 oracle;
    needs >= 114 frames for the guarded path to be taken."""
    from deeplip_b200.video_models import resnet as R
    from deeplip_b200.video_models.model import Lipreading
    sd = synth.make_video_state_dict(seed=seed, randomize=True)
    m = Lipreading(relu_type='prelu', backbone_type='resnet', extract_feats=True, tcn_options=synth.TCN_OPTIONS)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    raw = torch.from_numpy(synth.lip_crops_u8(list(range(B)), T=T, seed=seed))
    x = torch.stack([models_ref.video_preprocess(r) for r in raw])
    saved = R.USE_GUARDED
    with torch.no_grad():
        try:
            R.USE_GUARDED = True
            assert m.trunk.guarded_enabled(B * T, 11, 11)
            got = [m.trunk_maps(raw.to(DEV)).clone() for _ in range(reps)]
            feats = m(x[:, None].to(DEV), lengths=[T] * B)
            R.USE_GUARDED = False
            dense = m.trunk_maps(raw.to(DEV)).clone()
        finally:
            R.USE_GUARDED = saved
        ref = models_ref.lipreading_features(sd, x[:, None])                     # (B,T,512)
    torch.cuda.synchronize()
    out = {'mismatch_runs': sum(int(not torch.equal(got[0], g)) for g in got[1:]),
           'rel_vs_dense': rel_err(got[0].float().cpu(), dense.float().cpu()),
           'frame_cos_min': float(cosine_rows(feats.reshape(B * T, -1), ref.reshape(B * T, -1)).min())}
    assert out['mismatch_runs'] == 0 and out['rel_vs_dense'] < 1e-2 and out['frame_cos_min'] > 0.999, out
    return out


def video_golden_case():
    import os
    from deeplip_b200.video_models.model import Lipreading
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'video_small.npz'))
    sd = synth.make_video_state_dict(seed=1, randomize=True)
    m = Lipreading(relu_type='prelu', backbone_type='resnet', extract_feats=True, tcn_options=synth.TCN_OPTIONS)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    raw = torch.from_numpy(synth.lip_crops_u8([3, 3, 7], T=6, seed=1)).to(DEV)
    with torch.no_grad():
        y = m.trunk_maps(raw)
        feats, _ = ops.frame_pool_temporal_mean(y, 3, 6)
    torch.cuda.synchronize()
    ref = torch.from_numpy(gold['feats'])
    out = {'cos_min': float(cosine_rows(feats.reshape(18, -1), ref.reshape(18, -1)).min()), 'rel': rel_err(feats, ref)}
    assert out['cos_min'] > 0.999, out
    return out


def audio_model_case(arch='etdnn', pooling='statistic', B=3, nsamp=24000, seed=1, lengths=None):
    from deeplip_b200.audio_models.tdnn import SpeakerEmbNet
    o = synth.audio_opts(arch, pooling)
    sd = synth.make_audio_state_dict(o, seed=seed, randomize=True)
    net = SpeakerEmbNet(o)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    wav = synth.speech_like_audio(list(range(B)), nsamp=nsamp, seed=seed)
    feats = np.stack([frontend_np.extract_feature(w.astype(np.float64)).T for w in wav])     # (B,24,T)
    x = torch.from_numpy(feats)
    with torch.no_grad():
        xv_ref, xa_ref = models_ref.speaker_extract_embedding(sd, x, o)
        fw_ref = models_ref.speaker_forward(sd, x, o)
        xv, xa = net.extract_embedding(x.to(DEV))
        fw = net(x.to(DEV))
    torch.cuda.synchronize()
    out = {'xv_cos_min': float(cosine_rows(xv, xv_ref).min()), 'xa_cos_min': float(cosine_rows(xa, xa_ref).min()),
           'fw_cos_min': float(cosine_rows(fw, fw_ref).min()), 'xv_rel': rel_err(xv, xv_ref)}
    assert min(out['xv_cos_min'], out['xa_cos_min'], out['fw_cos_min']) > 0.999, out
    return out


def audio_golden_case():
    import os
    from deeplip_b200.audio_models.tdnn import SpeakerEmbNet
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'audio_small.npz'))
    wav = synth.speech_like_audio([3, 3, 7], nsamp=16000, seed=1)
    out = {}
    for arch in ('etdnn', 'tdnn'):
        for pool in ('statistic', 'attentive_statistic'):
            o = synth.audio_opts(arch, pool)
            net = SpeakerEmbNet(o)
            net.load_state_dict(synth.make_audio_state_dict(o, seed=1, randomize=True))
            net = net.to(DEV).eval()
            with torch.no_grad():
                f32, b16 = ops.frontend_features(torch.from_numpy(wav).to(DEV))
                xv, xa = net.embed_ntc(b16)
            torch.cuda.synchronize()
            k = '%s_%s' % (arch, pool)
            out[k] = float(cosine_rows(xv, torch.from_numpy(gold[k + '_xv'])).min())
            out[k + '_xa'] = float(cosine_rows(xa, torch.from_numpy(gold[k + '_xa'])).min())
    assert min(out.values()) > 0.999, out
    return out


def fusion_golden_case():
    import os
    from deeplip_b200.fusion_models.model_fusion import model_fusion, concat_fusion
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fusion_small.npz'))
    x = torch.from_numpy(synth.structured_embeddings([1, 1, 2, 3, 4], dim=1024, seed=1)).to(DEV)
    out = {}
    for ef in (True, False):
        m = model_fusion(1024, 512, 62, ef)
        m.load_state_dict(synth.make_fusion_state_dict(seed=1))
        m = m.to(DEV).eval()
        with torch.no_grad():
            y = m(x)
        torch.cuda.synchronize()
        out['linear_%d_cos' % ef] = float(cosine_rows(y, torch.from_numpy(gold['linear_%d' % ef])).min())
    c = concat_fusion(x[:, :512].contiguous(), x[:, 512:].contiguous())
    out['concat_abs'] = float((c.cpu() - torch.from_numpy(gold['concat'])).abs().max())
    assert out['linear_1_cos'] > 0.9999 and out['linear_0_cos'] > 0.9999 and out['concat_abs'] < 1e-5, out
    return out


make_trial_file = synth.make_trial_file


def scoring_full_case(tmpdir, kind='grid'):
    """Full 20 000-trial list: GPU scores vs the float64 oracle, EER within 0.05 % absolute,
    identical trial indexing, and size-independent properties (symmetry, self-score = 1)."""
    import os
    from deeplip_b200.trials import TrialList
    from deeplip_b200.fusion_models import utils as U
    if kind.endswith('_real'):      # the reference's own shipped lists (database/trial_*_v1.txt), reproduced under tests/golden
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'trial_%s_v1.txt' % kind[:-5])
    else:
        path = make_trial_file(os.path.join(tmpdir, 'trial_%s.txt' % kind), kind)
    tl = TrialList.from_file(path)
    if kind.endswith('_real'):
        assert len(tl) == 20000 and len(tl.utts) == {'grid_real': 25834, 'lomgrid_real': 3541}[kind]
    labels, pairs = scoring_ref.parse_trials(path)
    table, enrol, test = scoring_ref.utterance_table(pairs)
    assert table == tl.utts and np.array_equal(enrol, tl.enrol_idx) and np.array_equal(test, tl.test_idx)
    assert np.array_equal(labels, tl.labels)
    emb = synth.structured_embeddings([synth.speaker_of_utt(u) for u in tl.utts], dim=1024, seed=3, within=6.0)
    s = U.score_trials(emb, tl).cpu().numpy()
    ref = scoring_ref.cosine_scores_vec(emb, enrol, test)
    eer_g, thr_g = U.eer_from_scores(tl.labels, s)
    eer_r, thr_r = scoring_ref.eer_from_scores(labels, list(ref.astype(np.float32).reshape(-1, 1)))
    # properties: swapping enrol/test gives the same score; scoring an utterance against itself gives 1
    sw = ops.cosine_score_trials(torch.from_numpy(emb).to(DEV), torch.from_numpy(test).to(DEV),
                                 torch.from_numpy(enrol).to(DEV)).cpu().numpy()
    idx = torch.arange(len(tl.utts), dtype=torch.int32, device=DEV)
    self_s = ops.cosine_score_trials(torch.from_numpy(emb).to(DEV), idx, idx).cpu().numpy()
    dense = U.score_trials_dense(emb, tl).cpu().numpy()
    out = {'n_utts': len(tl.utts), 'score_abs': float(np.abs(s - ref).max()), 'eer': float(eer_g),
           'dense_abs': float(np.abs(dense - ref).max()),
           'eer_abs_diff': float(abs(eer_g - eer_r)), 'swap_abs': float(np.abs(s - sw).max()),
           'self_abs': float(np.abs(self_s - 1).max())}
    assert out['score_abs'] < 1e-5 and out['eer_abs_diff'] < 5e-4 and out['swap_abs'] == 0.0 and out['self_abs'] < 1e-6, out
    assert out['dense_abs'] < 1e-3, out          # bf16 operands: within north_star's per-trial tolerance
    assert 0.005 < out['eer'] < 0.45, out        # the synthetic list must not be degenerate
    return out


def pipeline_case(B=4, T=8, nsamp=24000, seed=1):
    """AV extraction end to end (wav + u8 crops -> fused embedding) vs the oracle chain; scores on all
    pairs within 1e-3; ragged batch == per-utterance runs."""
    from deeplip_b200.pipeline import AVExtractor, build_models
    audio, video = build_models(DEV, seed=seed)
    ex = AVExtractor(audio, video)
    spk = [1, 1, 2, 3][:B]
    wav = synth.speech_like_audio(spk, nsamp=nsamp, seed=seed)
    raw = synth.lip_crops_u8(spk, T=T, seed=seed)
    got = ex.extract(torch.from_numpy(wav).to(DEV), torch.from_numpy(raw).to(DEV))
    torch.cuda.synchronize()
    o = synth.audio_opts('etdnn', 'statistic')
    sda = synth.make_audio_state_dict(o, seed=seed)
    sdv = synth.make_video_state_dict(seed=seed)
    feats = torch.from_numpy(np.stack([frontend_np.extract_feature(w.astype(np.float64)).T for w in wav]))
    with torch.no_grad():
        xv, _ = models_ref.speaker_extract_embedding(sda, feats, o)
        x = torch.stack([models_ref.video_preprocess(torch.from_numpy(r)) for r in raw])
        em = models_ref.temporal_mean(models_ref.lipreading_features(sdv, x[:, None]))
        ref = models_ref.concat_fusion(xv, em)
    out = {'emb_cos_min': float(cosine_rows(got, ref).min())}
    pairs = [(i, j) for i in range(B) for j in range(B)]
    e = np.array([p[0] for p in pairs], dtype=np.int32)
    t = np.array([p[1] for p in pairs], dtype=np.int32)
    s_got = ops.cosine_score_trials(got, torch.from_numpy(e).to(DEV), torch.from_numpy(t).to(DEV)).cpu().numpy()
    s_ref = scoring_ref.cosine_scores_vec(ref.numpy(), e, t)
    out['score_abs'] = float(np.abs(s_got - s_ref).max())
    # ragged: utterance 1 truncated; batched (padded + lengths) must equal running it alone
    wl = torch.tensor([nsamp, nsamp - 5000] + [nsamp] * (B - 2), dtype=torch.int32, device=DEV)
    vl = torch.tensor([T, T - 3] + [T] * (B - 2), dtype=torch.int32, device=DEV)
    wav2 = wav.copy(); wav2[1, nsamp - 5000:] = 0
    raw_f = torch.stack([models_ref.video_preprocess(torch.from_numpy(r)) for r in raw])
    raw_f[1, T - 3:] = 0
    rag = ex.extract(torch.from_numpy(wav2).to(DEV), raw_f.to(DEV), wl, vl)
    alone = ex.extract(torch.from_numpy(wav2[1:2, :nsamp - 5000]).to(DEV), raw_f[1:2, :T - 3].contiguous().to(DEV))
    torch.cuda.synchronize()
    out['ragged_abs'] = float((rag[1] - alone[0]).abs().max())
    # host pipeline (pinned host buffers, H2D overlapped with compute) == direct device call
    from deeplip_b200.pipeline import HostPipeline
    hw, hv = torch.from_numpy(wav).pin_memory(), torch.from_numpy(raw).pin_memory()
    outs = HostPipeline(ex, DEV).run([(hw, hv), (hw, hv), (hw, hv)])
    out['host_pipeline_abs'] = max(float((o.to(DEV) - got).abs().max()) for o in outs)
    assert out['host_pipeline_abs'] == 0.0, out
    # ... and with lengths riding along (ragged batch from the host), alternating with a dense batch
    hw2, hv2 = torch.from_numpy(wav2).pin_memory(), raw_f.pin_memory()
    outs = HostPipeline(ex, DEV).run([(hw2, hv2, wl.cpu().pin_memory(), vl.cpu().pin_memory()), (hw, hv),
                                      (hw2, hv2, wl.cpu().pin_memory(), vl.cpu().pin_memory())])
    out['host_pipeline_ragged_abs'] = max(float((outs[0].to(DEV) - rag).abs().max()), float((outs[2].to(DEV) - rag).abs().max()),
                                          float((outs[1].to(DEV) - got).abs().max()))
    assert out['host_pipeline_ragged_abs'] == 0.0, out
    assert out['emb_cos_min'] > 0.999 and out['score_abs'] < 1e-3 and out['ragged_abs'] < 1e-5, out
    return out


def tcd_timit_ragged_case(durations=(3.0, 6.44, 10.0), seed=3):
    """BASELINE configs[4]: TCD-TIMIT-shaped long, variable-length utterances (3-10 s -> 75..250 lip frames at
    25 fps, 48 000..160 000 samples) in ONE padded batch with `lengths` (the reference's pad_packed_collate
    convention, models/video_models/dataset.py:123-139).  Each utterance of the ragged batch must equal the same
    utterance run alone, and the longest one must match the fp32 oracle chain."""
    from deeplip_b200.pipeline import AVExtractor, build_models
    audio, video = build_models(DEV, seed=seed)
    ex = AVExtractor(audio, video)
    B = len(durations)
    Tv = [int(round(25 * d)) for d in durations]
    ns = [int(16000 * d) for d in durations]
    Tm, nm = max(Tv), max(ns)
    spk = list(range(1, B + 1))
    wav = synth.speech_like_audio(spk, nsamp=nm, seed=seed)
    raw = synth.lip_crops_u8(spk, T=Tm, seed=seed)
    for i in range(B):                                   # zero-padded tails, like the collate function
        wav[i, ns[i]:] = 0
        raw[i, Tv[i]:] = 0
    wl = torch.tensor(ns, dtype=torch.int32, device=DEV)
    vl = torch.tensor(Tv, dtype=torch.int32, device=DEV)
    rag = ex.extract(torch.from_numpy(wav).to(DEV), torch.from_numpy(raw).to(DEV), wl, vl)
    out = {'alone_abs': 0.0}
    for i in range(B):
        alone = ex.extract(torch.from_numpy(wav[i:i + 1, :ns[i]].copy()).to(DEV),
                           torch.from_numpy(raw[i:i + 1, :Tv[i]].copy()).to(DEV))
        out['alone_abs'] = max(out['alone_abs'], float((rag[i] - alone[0]).abs().max()))
    torch.cuda.synchronize()
    i = int(np.argmax(Tv))
    o = synth.audio_opts('etdnn', 'statistic')
    sda, sdv = synth.make_audio_state_dict(o, seed=seed), synth.make_video_state_dict(seed=seed)
    with torch.no_grad():
        feat = torch.from_numpy(frontend_np.extract_feature(wav[i, :ns[i]].astype(np.float64)).T)[None]
        xv, _ = models_ref.speaker_extract_embedding(sda, feat, o)
        x = models_ref.video_preprocess(torch.from_numpy(raw[i, :Tv[i]]))[None, None]
        em = models_ref.temporal_mean(models_ref.lipreading_features(sdv, x))
        ref = models_ref.concat_fusion(xv, em)
    out['emb_cos_longest'] = float(cosine_rows(rag[i:i + 1], ref).min())
    assert out['alone_abs'] < 1e-4 and out['emb_cos_longest'] > 0.999, out
    return out


def graphed_extractor_case(B=4, T=8, nsamp=24000, seed=2):
    """GraphedExtractor: the whole step captured into one CUDA graph replays to the same bits as the eager launches,
    for the captured batch and for new data of the same shape."""
    from deeplip_b200.pipeline import AVExtractor, GraphedExtractor, build_models
    audio, video = build_models(DEV, seed=1)
    ex = AVExtractor(audio, video)
    spk = list(range(B))
    wav = torch.from_numpy(synth.speech_like_audio(spk, nsamp=nsamp, seed=seed)).to(DEV)
    raw = torch.from_numpy(synth.lip_crops_u8(spk, T=T, seed=seed)).to(DEV)
    wav2 = torch.from_numpy(synth.speech_like_audio(spk, nsamp=nsamp, seed=seed + 1)).to(DEV)
    raw2 = torch.from_numpy(synth.lip_crops_u8(spk, T=T, seed=seed + 1)).to(DEV)
    ref, ref2 = ex.extract(wav, raw).clone(), ex.extract(wav2, raw2).clone()
    g = GraphedExtractor(ex, wav, raw)
    out = {'same_batch': bool(torch.equal(g.extract(wav, raw), ref))}
    out['new_batch'] = bool(torch.equal(g.extract(wav2, raw2), ref2))
    torch.cuda.synchronize()
    try:
        g.extract(wav[:2], raw[:2])
        out['shape_check'] = False
    except RuntimeError:
        out['shape_check'] = True
    # Eager calls at OTHER shapes after capture (a dataset's tail batch) must not free or resize the persistent
    # buffers whose addresses the graph holds: run more shapes than ops.BUFFERS caches (forcing evictions), allocate
    # and scribble over fresh memory, then replay and compare bits.
    from deeplip_b200 import ops
    for b in range(1, B):
        for t in (T - 1, T + 1, T + 2):
            r = torch.from_numpy(synth.lip_crops_u8(spk[:b], T=t, seed=seed + 7)).to(DEV)
            ex.extract(wav[:b], r)
    junk = [torch.full((64 << 20,), 0x7f, dtype=torch.uint8, device=DEV) for _ in range(8)]
    torch.cuda.synchronize()
    out['replay_after_other_shapes'] = bool(torch.equal(g.extract(wav, raw), ref))
    out['replay_new_batch_after_other_shapes'] = bool(torch.equal(g.extract(wav2, raw2), ref2))
    del junk
    assert len(g._pinned_buffers) > 0
    assert all(out.values()), out
    return out


def full_size_batch_invariance_case(B=256, sub=64, seed=11):
    """BASELINE configs[2] size (batch 256 of GRID-shaped utterances: 75 x 96x96 u8 crops + 3 s of audio) through a
    size-independent property: every utterance's fused embedding in the batch of 256 is bit-identical to the one it
    gets in a batch of 64 (same kernels, same per-element summation order, no cross-utterance arithmetic), and two
    runs of the full batch are bit-identical."""
    import bench
    from deeplip_b200.pipeline import AVExtractor, build_models
    audio, video = build_models(DEV, seed=1)
    ex = AVExtractor(audio, video)
    raws, wavs = [], []
    for k in range(B // sub):
        r, w = bench.synth_batch(sub, seed=seed + k)
        raws.append(torch.from_numpy(r))
        wavs.append(torch.from_numpy(w))
    raw, wav = torch.cat(raws).to(DEV), torch.cat(wavs).to(DEV)
    big = ex.extract(wav, raw).clone()
    again = ex.extract(wav, raw).clone()
    parts = torch.cat([ex.extract(wav[i:i + sub].contiguous(), raw[i:i + sub].contiguous()).clone()
                       for i in range(0, B, sub)])
    torch.cuda.synchronize()
    out = {'rerun_equal': bool(torch.equal(big, again)), 'sub_batches_equal': bool(torch.equal(big, parts)),
           'finite': bool(torch.isfinite(big).all()), 'max_abs_diff': float((big - parts).abs().max())}
    assert out['rerun_equal'] and out['finite'] and out['max_abs_diff'] < 1e-5, out
    return out


def plda_case(dim=512, n_spk=40, per=25, n_trials=20000, seed=0):
    """PLDA trial scoring on a trial_grid-sized list against the oracle's per-trial loop
    (a sample of it: the loop is the reference's 2 590-trials/s path) and its vectorised closed form."""
    from deeplip_b200.plda import Classifier
    from oracle import plda_ref
    rng = np.random.default_rng(seed)
    spk = np.repeat(np.arange(n_spk), per)
    X = synth.structured_embeddings(spk.tolist(), dim=dim, seed=3).astype(np.float64)
    test_spk = np.repeat(np.arange(100, 133), 30)
    E = synth.structured_embeddings(test_spk.tolist(), dim=dim, seed=5).astype(np.float32)
    en = rng.integers(0, len(E), n_trials).astype(np.int32)
    te = rng.integers(0, len(E), n_trials).astype(np.int32)
    clf = Classifier().fit_model(X, spk, 20)
    got = clf.score_trials(torch.from_numpy(E).to(DEV), torch.from_numpy(en).to(DEV), torch.from_numpy(te).to(DEV))
    torch.cuda.synchronize()
    got = got.cpu().numpy().astype(np.float64)
    mo = plda_ref.fit(X, spk, 20)
    ref = plda_ref.plda_scores_loop(mo, E, en[:500], te[:500])
    scale = max(1.0, float(np.abs(ref).max()))
    out = {'llr_abs': float(np.abs(got[:500] - ref).max()), 'scale': scale}
    u = plda_ref.transform_D_to_U_model(mo, E)
    psi = mo['psi'][mo['relevant']]
    a, b = u[en], u[te]
    full = (np.log(psi + 1) - 0.5 * np.log(2 * psi + 1) + 0.5 * psi * (a + b) ** 2 / (2 * psi + 1)
            - 0.5 * psi * (a * a + b * b) / (psi + 1)).sum(1)
    out['llr_abs_all'] = float(np.abs(got - full).max())
    lab = (test_spk[en] == test_spk[te]).astype(int)
    e_got, e_ref = scoring_ref.eer_from_scores(lab, got)[0], scoring_ref.eer_from_scores(lab, full)[0]
    out['eer_abs'] = abs(float(e_got) - float(e_ref))
    assert out['llr_abs'] < 2e-3 * scale and out['llr_abs_all'] < 2e-3 * scale and out['eer_abs'] < 5e-4, out
    assert clf.score_trials(torch.from_numpy(E).to(DEV), torch.zeros(0, dtype=torch.int32, device=DEV),
                            torch.zeros(0, dtype=torch.int32, device=DEV)).numel() == 0
    return out


def determinism_pair_size_case(B=26, T=72, reps=30, seed=2):
    """The same at a size where every wide layer runs on the CTA-pair kernels with the staged (shared memory + TMA
    store) epilogue (1 872 frames): back-to-back launches, bitwise equal, and equal to the per-lane-store epilogue."""
    from deeplip_b200 import _lib
    from deeplip_b200.pipeline import AVExtractor, build_models
    audio, video = build_models(DEV, seed=seed)
    ex = AVExtractor(audio, video)
    spk = [1 + i % 7 for i in range(B)]
    raw = torch.from_numpy(synth.lip_crops_u8(spk, T=T, seed=seed)).to(DEV)
    wav = torch.from_numpy(synth.speech_like_audio(spk, nsamp=48000, seed=seed)).to(DEV)
    embs = [ex.extract(wav, raw).clone() for _ in range(reps)]
    maps = [video.trunk_maps(raw).clone() for _ in range(4)]
    _lib.set_option('staged_epilogue', 0)
    try:
        plain = ex.extract(wav, raw).clone()
    finally:
        _lib.set_option('staged_epilogue', 1)
    torch.cuda.synchronize()
    out = {'emb_mismatch': sum(int(not torch.equal(embs[0], e)) for e in embs[1:]),
           'maps_mismatch': sum(int(not torch.equal(maps[0], m)) for m in maps[1:]),
           'staged_equals_plain': bool(torch.equal(plain, embs[0]))}
    assert out['emb_mismatch'] == 0 and out['maps_mismatch'] == 0 and out['staged_equals_plain'], out
    return out


def determinism_case(B=8, T=10, reps=25, seed=1):
    """Bitwise run-to-run determinism of the video path under back-to-back launches (no host sync in
    between).  Guards the smem hand-offs between generic-proxy readers and TMA refills: a missing proxy
    fence showed up as ~5 % of launches with a few corrupted tiles, invisible to tolerance-based parity."""
    from deeplip_b200.pipeline import AVExtractor, build_models
    audio, video = build_models(DEV, seed=seed)
    ex = AVExtractor(audio, video)
    spk = list(range(B))
    raw = torch.from_numpy(synth.lip_crops_u8(spk, T=T, seed=seed)).to(DEV)
    wav = torch.from_numpy(synth.speech_like_audio(spk, nsamp=16000, seed=seed)).to(DEV)
    maps = [video.trunk_maps(raw).clone() for _ in range(reps)]
    embs = [ex.extract(wav, raw).clone() for _ in range(reps)]
    torch.cuda.synchronize()
    out = {'maps_mismatch': sum(int(not torch.equal(maps[0], m)) for m in maps[1:]),
           'emb_mismatch': sum(int(not torch.equal(embs[0], e)) for e in embs[1:])}
    assert out['maps_mismatch'] == 0 and out['emb_mismatch'] == 0, out
    return out


def video_tcn_case(B=2, T=12, seed=1):
    """Lipreading(extract_feats=False) drop-in (trunk + MS-TCN head) vs the oracle and the reference golden."""
    import os
    from deeplip_b200.video_models.model import Lipreading
    sd = synth.make_video_state_dict(seed=1, randomize=True)
    sd.update(synth.make_tcn_state_dict(num_classes=62, seed=1, randomize=True))
    m = Lipreading(relu_type='prelu', backbone_type='resnet', num_classes=62, extract_feats=False,
                   tcn_options=synth.TCN_OPTIONS)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    raw = torch.from_numpy(synth.lip_crops_u8([3, 7], T=12, seed=2))
    x = torch.stack([models_ref.video_preprocess(r) for r in raw])
    lengths = [12, 9]
    with torch.no_grad():
        ref = models_ref.ms_tcn_logits(sd, models_ref.lipreading_features(sd, x[:, None]), lengths)
        got = m(x[:, None].to(DEV), lengths=lengths)
    torch.cuda.synchronize()
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'video_tcn_small.npz'))['logits']
    out = {'cos_vs_oracle': float(cosine_rows(got, ref).min()), 'rel_vs_oracle': rel_err(got, ref),
           'cos_vs_reference_golden': float(cosine_rows(got, torch.from_numpy(gold)).min()),
           'argmax_equal': bool((got.argmax(1).cpu() == ref.argmax(1)).all())}
    assert out['cos_vs_oracle'] > 0.999 and out['cos_vs_reference_golden'] > 0.999 and out['argmax_equal'], out
    return out


def audio_resnet_case(B=3, Fd=24, T=120, pooling='average', seed=1):
    """Build-defined audio ResNet (absent upstream) vs the fp32 restatement of the same definition."""
    import copy
    from deeplip_b200.audio_models.resnet import SpeakerEmbNet
    opts = copy.deepcopy(synth.AUDIO_RESNET_OPTS)
    opts['resnet']['pooling'] = pooling
    sd = synth.make_audio_resnet_state_dict(opts, seed=seed)
    net = SpeakerEmbNet(opts)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    wav = synth.speech_like_audio(list(range(B)), nsamp=400 + 160 * (T - 1), seed=seed)
    feats = torch.from_numpy(np.stack([frontend_np.extract_feature(w.astype(np.float64)).T for w in wav]))[:, None]
    with torch.no_grad():
        ref, _ = models_ref.audio_resnet_extract_embedding(sd, feats, opts)
        got, _ = net.extract_embedding(feats.to(DEV))
    torch.cuda.synchronize()
    out = {'cos_min': float(cosine_rows(got, ref).min()), 'rel': rel_err(got, ref)}
    assert out['cos_min'] > 0.999, out
    return out


def trial_list_job_case(tmpdir, B=5, T=6, nsamp=16000, seed=3):
    """deeplip_b200.jobs.TrialListJob on one GPU with the real extractor: the table assembled batch by batch through
    `extract(out=table rows)` (K8 writing straight into the gather buffer) is bit-identical to the embeddings of ONE
    batched call, the scores equal the oracle's on that table, and the job's EER equals the oracle's EER."""
    from deeplip_b200.fusion_models import utils as U
    from deeplip_b200.jobs import TrialListJob
    from deeplip_b200.pipeline import AVExtractor, build_models
    from deeplip_b200.trials import TrialList
    audio, video = build_models(DEV, seed=1)
    ex = AVExtractor(audio, video)
    assert ex.dim == 1024
    tl = TrialList.from_file(make_trial_file(os.path.join(str(tmpdir), 'job.txt'), 'grid', n_target=30, n_non=60))
    n = len(tl.utts)
    spk = [synth.speaker_of_utt(u) for u in tl.utts]
    pool_spk = sorted(set(spk))
    wav_p = torch.from_numpy(synth.speech_like_audio(pool_spk, nsamp=nsamp, seed=seed, noise=0.3)).to(DEV)
    raw_p = torch.from_numpy(synth.lip_crops_u8(pool_spk, T=T, seed=seed, utt_sigma=1.0)).to(DEV)
    umap = torch.tensor([pool_spk.index(s) for s in spk], device=DEV)
    job = TrialListJob(tl, ex.dim, 0, 1, device=DEV, global_batch=B)

    def extract(lo, hi, out):
        idx = umap[lo:hi]
        got = ex.extract(wav_p.index_select(0, idx), raw_p.index_select(0, idx), out=out)
        assert got.data_ptr() == out.data_ptr()

    res = job.run(extract, lambda tab, en, te: ops.cosine_score_trials(tab, en, te), eer_fn=U.eer_from_scores)
    whole = ex.extract(wav_p.index_select(0, umap), raw_p.index_select(0, umap))
    torch.cuda.synchronize()
    assert job.table.shape == (n, 1024) and torch.equal(job.table, whole), 'in-table extraction differs from one batched call'
    ref = scoring_ref.cosine_scores_vec(whole.cpu().numpy(), tl.enrol_idx, tl.test_idx)
    err = float(np.abs(res['scores'].cpu().numpy() - ref).max())
    assert err < 1e-5, err
    ref_eer, _ = scoring_ref.eer_from_scores(tl.labels, list(ref.astype(np.float32).reshape(-1, 1)))
    assert abs(res['eer'] - ref_eer) < 5e-4
    assert job.verify_gather() and set(res['ms']) == {'extract', 'checksum', 'rank_skew', 'all_gather', 'score', 'gather_scores'}
    return {'n_utts': n, 'score_abs': err, 'eer': float(res['eer'])}


def frontend_pcm16_case(B=3, nsamp=24000, seed=4):
    """int16 PCM input (dl_frontend_features_pcm16) == the f32 entry point fed value / 32768, bit for bit, for every
    feature kind; and AVExtractor takes either."""
    rng = np.random.default_rng(seed)
    pcm = torch.from_numpy(np.clip(np.rint(synth.speech_like_audio(list(range(B)), nsamp=nsamp, seed=seed) * 32767), -32768,
                                   32767).astype(np.int16)).to(DEV)
    lengths = torch.tensor([nsamp, nsamp - 3001, nsamp - 77][:B], dtype=torch.int32, device=DEV)
    out = {}
    for kind, nf in (('mfcc', 24), ('fbank', 24), ('logfbank', 60), ('stft', 257)):
        for ln in (None, lengths):
            a32, ab = ops.frontend_features(pcm, kind, nf, True, lengths=ln)
            b32, bb = ops.frontend_features(pcm.float() / 32768.0, kind, nf, True, lengths=ln)
            torch.cuda.synchronize()
            out['%s%s' % (kind, '' if ln is None else '_ragged')] = bool(torch.equal(a32, b32) and torch.equal(ab, bb))
    assert all(out.values()), out
    return out


def avgpool_fused_case(N=2400, C=512, lengths=(75, 3, 40), seed=0):
    """K4 in the epilogue of a 3x3 conv with residual (layer4's last conv at pair-kernel size): the image means taken
    from the staged output tiles (126-row m tiles) equal, BIT FOR BIT, the unfused path (conv as usual, then the pooling
    kernel), with and without the bf16 tensor being written; and the temporal half run on those frame features equals
    the one-kernel K4 (ragged lengths included)."""
    from deeplip_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(N, 3, 3, C, device=DEV, generator=g).to(torch.bfloat16)
    res = torch.randn(N, 3, 3, C, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(C, 9 * C, device=DEV, generator=g) * 0.02).to(torch.bfloat16)
    sc = torch.rand(C, device=DEV, generator=g) + 0.5
    sh = torch.randn(C, device=DEV, generator=g) * 0.1
    sl = torch.rand(C, device=DEV, generator=g) * 0.5
    conv = lambda **kw: ops.conv_igemm(x, w, C, C, 3, 3, (1, 1), (1, 1), (1, 1), sc, sh, sl, residual=res, **kw)
    y_ref, _ = conv()
    out = {}
    _lib.set_option('pool_fuse', 0)
    try:
        y0, p0 = conv(avgpool=True, avgpool_keep_y=True)
    finally:
        _lib.set_option('pool_fuse', 1)
    y1, p1 = conv(avgpool=True, avgpool_keep_y=True)
    _, p2 = conv(avgpool=True)
    torch.cuda.synchronize()
    out['y_unfused_equal'] = bool(torch.equal(y0, y_ref))
    out['y_fused_equal'] = bool(torch.equal(y1, y_ref))
    out['pool_fused_equals_unfused'] = bool(torch.equal(p0, p1))
    out['pool_without_y_equal'] = bool(torch.equal(p0, p2))
    ref = y_ref.float().view(N, 9, C).mean(1)
    out['pool_abs_vs_torch'] = float((p1 - ref).abs().max())
    T = 75
    B = N // T
    ln = torch.tensor([lengths[i % len(lengths)] for i in range(B)], dtype=torch.int32, device=DEV)
    ff, um = ops.frame_pool_temporal_mean(y_ref, B, T, lengths=ln)
    um2 = ops.temporal_mean(p1, B, T, lengths=ln)
    torch.cuda.synchronize()
    out['frame_feats_equal'] = bool(torch.equal(ff.view(N, C), p1))
    out['temporal_mean_equal'] = bool(torch.equal(um, um2))
    assert all(v for k, v in out.items() if k != 'pool_abs_vs_torch') and out['pool_abs_vs_torch'] < 1e-4, out
    return out


def avgpool_model_case(B=28, T=75, seed=3):
    """Lipreading.utterance_embedding at a batch that puts layer4's last conv on the pair kernel (fused K4) against the
    same call with the fusion switched off, and against every clip run alone (small batches take the unfused path):
    bit-identical, i.e. an utterance's embedding does not depend on which path its batch size selects."""
    from deeplip_b200 import _lib
    from deeplip_b200.pipeline import build_models
    _, video = build_models(DEV, seed=seed)
    raw = torch.from_numpy(synth.lip_crops_u8(list(range(B)), T=T, H=96, W=96, seed=seed)).to(DEV)
    ln = torch.tensor([T - (3 * i) % 40 for i in range(B)], dtype=torch.int32, device=DEV)
    fused = video.utterance_embedding(raw, ln).clone()
    _lib.set_option('pool_fuse', 0)
    try:
        plain = video.utterance_embedding(raw, ln).clone()
    finally:
        _lib.set_option('pool_fuse', 1)
    alone = torch.cat([video.utterance_embedding(raw[i:i + 1, :int(ln[i])].contiguous()) for i in (0, 5, B - 1)])
    torch.cuda.synchronize()
    out = {'fused_equals_unfused': bool(torch.equal(fused, plain)),
           'batch_equals_alone': bool(torch.equal(fused[[0, 5, B - 1]], alone))}
    assert all(out.values()), out
    return out


def prepass_overlap_case(B=3, T=7, seed=2):
    """AVExtractor.overlap_prepass: the stem's pre-pass on a side stream under the audio branch (dl_stem_prepass +
    dl_stem_conv3d_prepassed) gives the bits of the one-call form, ragged lengths included, also back to back."""
    from deeplip_b200.pipeline import AVExtractor, build_models
    audio, video = build_models(DEV, seed=seed)
    ex = AVExtractor(audio, video)
    raw = torch.from_numpy(synth.lip_crops_u8(list(range(B)), T=T, H=96, W=96, seed=seed)).to(DEV)
    wav = torch.from_numpy(synth.speech_like_audio(list(range(B)), nsamp=16000, seed=seed)).to(DEV)
    vl = torch.tensor([T, T - 3, T - 1][:B], dtype=torch.int32, device=DEV)
    ref = [ex.extract(wav, raw).clone(), ex.extract(wav, raw, None, vl).clone()]
    ex.overlap_prepass = True
    got = [ex.extract(wav, raw).clone(), ex.extract(wav, raw, None, vl).clone()]
    rep = [ex.extract(wav, raw).clone() for _ in range(4)]
    torch.cuda.synchronize()
    out = {'equal': bool(torch.equal(ref[0], got[0])), 'ragged_equal': bool(torch.equal(ref[1], got[1])),
           'back_to_back_equal': all(bool(torch.equal(r, ref[0])) for r in rep)}
    assert all(out.values()), out
    return out
