"""ctypes binding of libdeeplip_b200.so (the C ABI declared in include/deeplip_b200.h).

There is NO fallback: if the shared library is missing, or a call returns a non-zero status
(e.g. no sm_100 device), a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdeeplip_b200.so')

DL_OK = 0


class ConvDesc(C.Structure):
    """struct dl_conv_desc (include/deeplip_b200.h)."""
    _fields_ = [(n, C.c_int) for n in
                ('N', 'H', 'W', 'C', 'ldx', 'Cout', 'R', 'S', 'stride_h', 'stride_w', 'pad_h', 'pad_w',
                 'dil_h', 'dil_w', 'ldy', 'ldf')] + [('f32_slope', C.c_float)] + [(n, C.c_int) for n in
                ('img_rows', 'img_cols', 'lin', 'valid_h', 'valid_w', 'out_img_rows', 'out_img_cols',
                 'split_channel', 'split_center_only')] + [('y_split', C.c_void_p), ('center_only_from', C.c_int),
                                                                      ('avgpool', C.c_int), ('avgpool_keep_y', C.c_int), ('avgpool_out', C.c_void_p)]


_p, _i, _f = C.c_void_p, C.c_int, C.c_float

# name -> argtypes ; every compute entry point returns int
SIGNATURES = {
    'dl_frontend_features': [_p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _i, _p],
    'dl_frontend_features_pcm16': [_p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _i, _p],
    'dl_nct_to_ntc_bf16': [_p, _i, _i, _i, _p, _i, _p],
    'dl_stem_conv3d_bn_prelu_pool': [_p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p, _i, _p, _p, _p],
    'dl_stem_prepass': [_p, _i, _i, _i, _i, _i, _i, _i, _f, _f, _p, _p, _p],
    'dl_stem_conv3d_prepassed': [_i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p],
    'dl_conv3x3_c64_halo_bf16': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    'dl_conv_igemm_bf16': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, C.POINTER(ConvDesc), _p],
    'dl_frame_pool_temporal_mean': [_p, _i, _i, _i, _i, _p, _p, _p, _p],
    'dl_temporal_mean_f32': [_p, _i, _i, _i, _p, _p, _p],
    'dl_stat_pool': [_p, _i, _i, _i, _i, _p, _p, _p, _i, _p],
    'dl_attn_stat_pool': [_p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p],
    'dl_attn_logits': [_p, _i, _i, _i, _p, _f, _p, _p],
    'dl_znorm_concat': [_p, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p],
    'dl_lowfer': [_p, _p, _i, _i, _p, _p],
    'dl_l2_normalize': [_p, _i, _i, _p, _p, _p],
    'dl_affine_act': [_p, _i, _i, _p, _p, _f, _p, _i, _p, _p],
    'dl_cosine_score_trials': [_p, _i, _i, _p, _p, _i, _p, _p],
    'dl_score_fusion_trials': [_p, _i, _p, _i, _i, _p, _p, _i, _p, _p],
    'dl_gather_scores': [_p, _i, _p, _p, _i, _p, _p],
    'dl_plda_transform': [_p, _i, _i, _p, _p, _i, _p, _p],
    'dl_plda_llr_trials': [_p, _i, _i, _p, _p, _f, _p, _p, _i, _p, _p],
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'deeplip_b200: %s not found -- build it with `python -m deeplip_b200.build` '
                '(there is no CPU / PyTorch fallback)' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        l.dl_version.restype = C.c_int
        l.dl_last_error.restype = C.c_char_p
        l.dl_launch_count.restype = C.c_longlong
        l.dl_set_option.argtypes = [C.c_char_p, _i]
        l.dl_set_option.restype = C.c_int
        l.dl_stem_workspace_bytes.restype = C.c_longlong
        l.dl_stem_workspace_bytes.argtypes = [_i, _i, _i, _i]
        _lib = l
    return _lib


def check(status, what):
    if status != DL_OK:
        raise RuntimeError('deeplip_b200.%s failed (%d): %s' % (what, status, lib().dl_last_error().decode()))


def set_option(name, value):
    check(lib().dl_set_option(name.encode(), int(value)), 'dl_set_option')


def launch_count():
    return int(lib().dl_launch_count())
