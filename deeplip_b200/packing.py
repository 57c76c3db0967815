"""Weight repacking for the sm_100a kernels (done once per load_state_dict, on the device).

BatchNorm (eval) folds into a per-channel (scale, shift) that the kernels apply in fp32 in the
epilogue -- weights themselves are only cast to bf16, never pre-multiplied, so the rounding of
each weight is independent of the BN statistics.
"""
import torch

BN_EPS = 1e-5


def ceil_to(x, m):
    return (x + m - 1) // m * m


def fold_bn(bn_weight, bn_bias, mean, var, conv_bias=None, eps=BN_EPS):
    """y = bn(conv + bias) = conv * scale + shift."""
    scale = bn_weight.float() / torch.sqrt(var.float() + eps)
    shift = bn_bias.float() - mean.float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.float() * scale
    return scale.contiguous(), shift.contiguous()


def pad_vec(v, n, value=0.0):
    out = torch.full((n,), value, device=v.device, dtype=torch.float32)
    out[:v.numel()] = v.float()
    return out


def pack_conv_weight(w, cout_pad=None):
    """(Cout, Cin, R, S) -> (Cout_pad, R*S*ceil64(Cin)) bf16, K index = (r*S+s)*ceil64(Cin)+c
    (the layout dl_conv_igemm_bf16 documents)."""
    Cout, Cin, R, S = w.shape
    cpad = ceil_to(Cin, 64)
    cout_pad = cout_pad or ceil_to(Cout, 8)
    out = torch.zeros((cout_pad, R * S, cpad), device=w.device, dtype=torch.float32)
    out[:Cout, :, :Cin] = w.float().permute(0, 2, 3, 1).reshape(Cout, R * S, Cin)
    return out.reshape(cout_pad, R * S * cpad).to(torch.bfloat16).contiguous()


def pack_conv1d_weight(w, cout_pad=None):
    """(Cout, Cin, k) Conv1d weight -> conv2d with R=1, S=k."""
    return pack_conv_weight(w.unsqueeze(2), cout_pad)


def pack_linear_weight(w, cout_pad=None):
    return pack_conv_weight(w[:, :, None, None], cout_pad)


def pack_stem_weight(w):
    """(64, 1, 5, 7, 7) Conv3d weight -> (64, 320) bf16, K = kt*64 + kh*8 + kw, zero padded
    (dl_stem_conv3d_bn_prelu_pool)."""
    assert tuple(w.shape) == (64, 1, 5, 7, 7)
    out = torch.zeros((64, 5, 8, 8), device=w.device, dtype=torch.float32)
    out[:, :, :7, :7] = w.float()[:, 0]
    return out.reshape(64, 320).to(torch.bfloat16).contiguous()
