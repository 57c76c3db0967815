"""Torch-tensor front door of the C ABI: pointer / shape plumbing only (no math happens here).

Every function takes CUDA tensors, enqueues one or two kernels of libdeeplip_b200.so on the
current CUDA stream and returns freshly allocated outputs.  bf16 activations are channels-last.
"""
import ctypes as C
import torch

from . import _lib
from ._lib import ConvDesc


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('deeplip_b200 ops need CUDA tensors (there is no CPU fallback)')


class BufferCache:
    """Persistent device buffers keyed by shape (stem workspace, the trunk's zero-padded activation layouts).

    A buffer is never resized in place: a new shape gets a new buffer and the old one stays cached (LRU, `cap`
    shapes), so raw pointers that an earlier call baked into a CUDA graph stay valid while the buffer is cached;
    GraphedExtractor additionally pins what it captured with `recording()`, which keeps those tensors alive after an
    eviction.  One model instance drives ONE stream at a time: two streams through the same cached buffers race
    (as they would through the reference's in-place BatchNorm buffers)."""

    def __init__(self, cap=8):
        from collections import OrderedDict
        self._d = OrderedDict()
        self._cap = cap
        self._rec = None

    def get(self, key, make):
        buf = self._d.get(key)
        if buf is None:
            buf = make()
            self._d[key] = buf
            while len(self._d) > self._cap:
                self._d.popitem(last=False)
        else:
            self._d.move_to_end(key)
        if self._rec is not None:
            self._rec.append(buf)
        return buf

    def recording(self):
        """Context manager: collects every buffer handed out inside it (returned list keeps them alive)."""
        cache = self

        class _Rec:
            def __enter__(self):
                self.prev, cache._rec = cache._rec, []
                return cache._rec

            def __exit__(self, *exc):
                got, cache._rec = cache._rec, self.prev
                if self.prev is not None:
                    self.prev.extend(got)
                return False
        return _Rec()

    def clear(self):
        self._d.clear()


BUFFERS = BufferCache()


def _workspace(device, nbytes):
    """Per-device scratch the C ABI asks the caller to own, one buffer per size (see BufferCache)."""
    return BUFFERS.get(('ws', str(device), int(nbytes)),
                       lambda: torch.empty(int(nbytes), dtype=torch.uint8, device=device))


def ceil_to(x, m):
    return (x + m - 1) // m * m


def num_frames(nsamp, frame_len=400, step=160, feat_type='mfcc'):
    """Frames of `nsamp` samples: python_speech_features framing (ceil, tail zero-padded), or for 'stft'
    librosa's centre-padded framing (1 + nsamp // hop)."""
    if feat_type == 'stft':
        return 1 + nsamp // step
    return 1 if nsamp <= frame_len else 1 + -(-(nsamp - frame_len) // step)


# ------------------------------------------------------------------ K1
FEAT_KINDS = {'mfcc': 0, 'fbank': 1, 'logfbank': 2, 'stft': 3}


def frontend_features(wav, feat_type='mfcc', n_feat=24, cmvn=True, lengths=None, ld=None, delta=False):
    """wav (B, nsamp) f32, or int16 PCM (sample = value / 32768, as soundfile.read decodes a 16-bit file; same bits
    out) -> (feat_f32 (B,F,T), feat_bf16 (B,T,ld)).  feat_type 'stft' (n_fft 512, hop 160,
    Hann 400; models/fusion_models/datasets.py:237-241) has n_feat = 257.  delta (True = the reference's order 2, or
    0 / 1 / 2): `_delta` (:217-225) appends delta(feat, 1) and delta(feat, 2): F becomes n_feat * (1 + order)."""
    _need_cuda(wav, lengths)
    if feat_type not in FEAT_KINDS:
        raise NotImplementedError('Other features are not implemented!')      # datasets.py:243
    pcm16 = wav.dtype == torch.int16
    wav = wav.contiguous() if pcm16 else wav.contiguous().float()
    B, nsamp = wav.shape
    if feat_type == 'stft':
        n_feat = 257
    T = num_frames(nsamp, feat_type=feat_type)
    order = 2 if delta is True else int(delta)
    n_out = n_feat * (1 + order)
    ld = ld or ceil_to(n_out, 64)
    f32 = torch.empty((B, n_out, T), device=wav.device, dtype=torch.float32)
    b16 = torch.empty((B, T, ld), device=wav.device, dtype=torch.bfloat16)
    fn = _lib.lib().dl_frontend_features_pcm16 if pcm16 else _lib.lib().dl_frontend_features
    st = fn(_ptr(wav), _ptr(lengths), B, nsamp, FEAT_KINDS[feat_type], n_feat,
            int(bool(cmvn)), order, _ptr(b16), ld, _ptr(f32), T, _stream())
    _lib.check(st, 'dl_frontend_features')
    return f32, b16


def nct_to_ntc_bf16(x, ld=None):
    _need_cuda(x)
    x = x.contiguous().float()
    B, Cc, T = x.shape
    ld = ld or ceil_to(Cc, 64)
    y = torch.empty((B, T, ld), device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().dl_nct_to_ntc_bf16(_ptr(x), B, Cc, T, _ptr(y), ld, _stream()), 'dl_nct_to_ntc_bf16')
    return y


# ------------------------------------------------------------------ K2
def stem_conv3d(x, w_packed, scale, shift, slope, crop=(88, 88), mean=0.421, std=0.165, out=None, lengths=None,
                phase='both'):
    """x: (B,T,H,W) f32 normalised frames, or (B,T,Hraw,Wraw) uint8 raw crops -> (B*T, H/4, W/4, 64) bf16.
    out: optional pre-zeroed (B*T, rows >= H/4, W/4, 64) buffer in the stacked-rows layout.
    lengths: int32 CUDA (B,) valid frames per clip; later frames are treated as zero normalised frames.
    phase: 'both' (default), or the two halves as calls of their own -- 'prepass' (x -> the cached workspace; returns
    None; w_packed / scale / shift / slope may be None) and 'main' (workspace -> y; x is only consulted for its shape)
    -- for callers that run the pre-pass on another stream (AVExtractor does, under the audio branch)."""
    _need_cuda(x, w_packed, scale, shift, slope, lengths)
    if lengths is not None:
        assert lengths.dtype == torch.int32 and lengths.numel() == x.shape[0]
    x = x.contiguous()
    B, T = x.shape[0], x.shape[1]
    if x.dtype == torch.uint8:
        Hraw, Wraw = x.shape[2], x.shape[3]
        H, W = crop
        is_u8 = 1
    else:
        x = x.float()
        H, W = x.shape[2], x.shape[3]
        Hraw, Wraw, is_u8 = H, W, 0
    ws = _workspace(x.device, _lib.lib().dl_stem_workspace_bytes(B, T, H, W))
    if phase == 'prepass':
        st = _lib.lib().dl_stem_prepass(_ptr(x), is_u8, B, T, H, W, Hraw, Wraw, float(mean), float(std), _ptr(lengths),
                                        _ptr(ws), _stream())
        _lib.check(st, 'dl_stem_prepass')
        return None
    if out is None:
        y = torch.empty((B * T, H // 4, W // 4, 64), device=x.device, dtype=torch.bfloat16)
    else:
        y = out
        assert y.dtype == torch.bfloat16 and y.is_contiguous() and y.shape[0] == B * T and \
            y.shape[1] >= H // 4 and y.shape[2] == W // 4 and y.shape[3] == 64
    if phase == 'main':
        st = _lib.lib().dl_stem_conv3d_prepassed(B, T, H, W, _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope), _ptr(y),
                                                 y.shape[1], _ptr(ws), _stream())
        _lib.check(st, 'dl_stem_conv3d_prepassed')
        return y
    assert phase == 'both'
    st = _lib.lib().dl_stem_conv3d_bn_prelu_pool(_ptr(x), is_u8, B, T, H, W, Hraw, Wraw, float(mean), float(std),
                                                 _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope), _ptr(y),
                                                 y.shape[1], _ptr(lengths), _ptr(ws), _stream())
    _lib.check(st, 'dl_stem_conv3d_bn_prelu_pool')
    return y


# ------------------------------------------------------------------ K3 / K5 / K7
def conv_igemm(x, w_packed, Cin, Cout, R=1, S=1, stride=(1, 1), pad=(0, 0), dil=(1, 1), scale=None, shift=None,
               slope=None, residual=None, want_bf16=True, want_f32=False, scale2=None, shift2=None,
               f32_slope=1.0, H=None, W=None, out=None, out_channel_offset=0, residual_channel_offset=0, split=None,
               center_only_from=0, avgpool=False, avgpool_keep_y=False):
    """x: (N,H,W,ldx) bf16 channels-last -- or (N,img_rows,img_cols,ldx) stacked / guarded with the true extents
    passed as H, W.  `out` may be a wider (channel slice) and / or guarded (N, >=P, >=Q, ld) caller-owned buffer;
    with `out`, `residual` is a buffer of out's shape read at `residual_channel_offset` (the kernel indexes the residual
    with the output's pixel pitch).
    split=(c, center_only): sibling convs in one launch -- output channels >= c go to a second dense tensor and, with
    center_only, are declared to have zero weights off the centre tap (dl_conv_desc.split_channel); the first return
    value is then the pair (y[..., :c], y[..., c:]) as two dense tensors.
    center_only_from=c: declares channels >= c centre-tap-only without splitting the output (dl_conv_desc.center_only_from).
    avgpool=True (K4 in the epilogue, dl_conv_desc.avgpool): the second return value is the global average pool of every
    output image, (N, Cout) f32; the first one is only defined with avgpool_keep_y (else it is scratch the fused path
    does not write).
    Returns (y_bf16 (N,P,Q,Cout) | out | None, y_f32 (N*P*Q,Cout) | None)."""
    _need_cuda(x, w_packed, scale, shift, slope, residual, scale2, shift2)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    N, img_rows, img_cols, ldx = x.shape
    H = H or img_rows
    W = W or img_cols
    P = (H + 2 * pad[0] - dil[0] * (R - 1) - 1) // stride[0] + 1
    Q = (W + 2 * pad[1] - dil[1] * (S - 1) - 1) // stride[1] + 1
    ldy = Cout
    y_ptr = None
    out_rows = out_cols = 0
    if split is not None:
        sc, center_only = split
        assert out is None and residual is None and not want_f32 and 0 < sc < Cout
        ya = torch.empty((N, P, Q, sc), device=x.device, dtype=torch.bfloat16)
        yb = torch.empty((N, P, Q, Cout - sc), device=x.device, dtype=torch.bfloat16)
        ldy = max(sc, Cout - sc)
        assert sc == Cout - sc, 'both halves share one pitch'
        d = ConvDesc(N, H, W, Cin, ldx, Cout, R, S, stride[0], stride[1], pad[0], pad[1], dil[0], dil[1], ldy, Cout,
                     float(f32_slope), img_rows, img_cols, 0, 0, 0, 0, 0, sc, int(bool(center_only)), yb.data_ptr())
        st = _lib.lib().dl_conv_igemm_bf16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope), None,
                                           _ptr(ya), None, None, None, C.byref(d), _stream())
        _lib.check(st, 'dl_conv_igemm_bf16')
        return (ya, yb), None
    if out is not None:          # write Cout channels into a slice of a wider channels-last buffer (concat for free)
        assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.dim() == 4 and out.shape[0] == N
        assert out.shape[1] >= P and out.shape[2] >= Q
        assert out_channel_offset % 8 == 0 and out_channel_offset + Cout <= out.shape[3]
        if tuple(out.shape[1:3]) != (P, Q):
            out_rows, out_cols = out.shape[1], out.shape[2]
            assert not want_f32
        y, ldy = out, out.shape[3]
        y_ptr = C.c_void_p(out.data_ptr() + 2 * out_channel_offset)
        assert residual is None or (out_channel_offset == 0 and residual.shape == out.shape and
                                    residual_channel_offset % 8 == 0 and
                                    residual_channel_offset + Cout <= residual.shape[3])
    else:
        y = torch.empty((N, P, Q, Cout), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    yf = torch.empty((N * P * Q, Cout), device=x.device, dtype=torch.float32) if want_f32 else None
    r_ptr = _ptr(residual)
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous() and residual.numel() == y.numel()
        assert residual_channel_offset == 0 or out is not None
        r_ptr = C.c_void_p(residual.data_ptr() + 2 * residual_channel_offset)
    d = ConvDesc(N, H, W, Cin, ldx, Cout, R, S, stride[0], stride[1], pad[0], pad[1], dil[0], dil[1], ldy, Cout,
                 float(f32_slope), img_rows, img_cols, 0, 0, 0, out_rows, out_cols, 0, 0, None, int(center_only_from))
    if avgpool:
        assert out is None and not want_f32 and want_bf16
        yf = torch.empty((N, Cout), device=x.device, dtype=torch.float32)      # image means ride in the f32 slot of the return
        d.avgpool, d.avgpool_keep_y, d.avgpool_out = 1, int(bool(avgpool_keep_y)), yf.data_ptr()
        st = _lib.lib().dl_conv_igemm_bf16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope),
                                           r_ptr, _ptr(y), None, _ptr(scale2), _ptr(shift2), C.byref(d), _stream())
        _lib.check(st, 'dl_conv_igemm_bf16')
        return (y if avgpool_keep_y else None), yf
    st = _lib.lib().dl_conv_igemm_bf16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope),
                                       r_ptr, y_ptr if y_ptr is not None else _ptr(y), _ptr(yf),
                                       _ptr(scale2), _ptr(shift2), C.byref(d), _stream())
    _lib.check(st, 'dl_conv_igemm_bf16')
    return y, yf


def conv_igemm_lin(x, w_packed, Cin, Cout, valid, R=1, S=1, pad=(0, 0), dil=(1, 1), scale=None, shift=None,
                   slope=None, residual=None, out=None):
    """Stride-1 convolution on a guarded tensor (include/deeplip_b200.h, "Guarded layouts"): x, out and residual
    are (N, Hg, Wg, ld) bf16 whose guard rows / columns (>= pad below / right of every image) are zero; outputs
    are stored for h < valid[0], w < valid[1] only.  `out` None allocates a zeroed tensor."""
    _need_cuda(x, w_packed, scale, shift, slope, residual, out)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 4
    N, Hg, Wg, ldx = x.shape
    if out is None:
        out = torch.zeros((N, Hg, Wg, Cout), device=x.device, dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and tuple(out.shape[:3]) == (N, Hg, Wg)
    assert out.shape[3] >= Cout
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous() and residual.shape == out.shape
    d = ConvDesc(N, Hg, Wg, Cin, ldx, Cout, R, S, 1, 1, pad[0], pad[1], dil[0], dil[1], out.shape[3], Cout,
                 1.0, 0, 0, 1, valid[0], valid[1], 0, 0)
    st = _lib.lib().dl_conv_igemm_bf16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope),
                                       _ptr(residual), _ptr(out), None, None, None, C.byref(d), _stream())
    _lib.check(st, 'dl_conv_igemm_bf16')
    return out


def conv3x3_halo(x, w_packed, scale, shift, slope, H, out, residual=None):
    """3x3 s1 p1 64->64 conv on stacked-rows activations: x, out, residual (N, img_rows >= H+1, W, 64) bf16 with
    zero padding rows; `out` is a caller-owned buffer whose padding rows are already zero."""
    _need_cuda(x, w_packed, scale, shift, slope, residual, out)
    for t in (x, out, residual):
        assert t is None or (t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape == x.shape)
    N, img_rows, W, Cc = x.shape
    assert Cc == 64
    st = _lib.lib().dl_conv3x3_c64_halo_bf16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(shift), _ptr(slope),
                                             _ptr(residual), _ptr(out), N, H, W, img_rows, _stream())
    _lib.check(st, 'dl_conv3x3_c64_halo_bf16')
    return out


# ------------------------------------------------------------------ K4 / K6
def frame_pool_temporal_mean(x, B, T, lengths=None, want_frames=True, want_mean=True):
    """x: (B*T, P, Q, C) bf16 -> (frame_feats (B,T,C) f32 | None, utt_mean (B,C) f32 | None)."""
    _need_cuda(x, lengths)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    Cc = x.shape[-1]
    HW = x.numel() // (B * T * Cc)
    ff = torch.empty((B, T, Cc), device=x.device, dtype=torch.float32) if want_frames else None
    um = torch.empty((B, Cc), device=x.device, dtype=torch.float32) if want_mean else None
    st = _lib.lib().dl_frame_pool_temporal_mean(_ptr(x), B, T, HW, Cc, _ptr(lengths), _ptr(ff), _ptr(um), _stream())
    _lib.check(st, 'dl_frame_pool_temporal_mean')
    return ff, um


def temporal_mean(frame_feats, B, T, lengths=None):
    """frame_feats: (B*T, C) or (B, T, C) f32 -> (B, C) f32 mean over the first lengths[b] frames (the temporal half of
    K4; same bits as frame_pool_temporal_mean's utt_mean)."""
    _need_cuda(frame_feats, lengths)
    assert frame_feats.dtype == torch.float32 and frame_feats.is_contiguous()
    Cc = frame_feats.shape[-1]
    assert frame_feats.numel() == B * T * Cc
    um = torch.empty((B, Cc), device=frame_feats.device, dtype=torch.float32)
    st = _lib.lib().dl_temporal_mean_f32(_ptr(frame_feats), B, T, Cc, _ptr(lengths), _ptr(um), _stream())
    _lib.check(st, 'dl_temporal_mean_f32')
    return um


def stat_pool(x, Cc, lengths=None, want_f32=True, want_bf16=True, logits=None):
    """x: (B,T,ldx) bf16 -> (out_f32 (B,2C) | None, out_bf16 (B,2C) | None); attentive if logits given."""
    _need_cuda(x, lengths, logits)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.dim() == 3
    B, T, ldx = x.shape
    of = torch.empty((B, 2 * Cc), device=x.device, dtype=torch.float32) if want_f32 else None
    ob = torch.empty((B, 2 * Cc), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    if logits is None:
        st = _lib.lib().dl_stat_pool(_ptr(x), B, T, Cc, ldx, _ptr(lengths), _ptr(of), _ptr(ob), 2 * Cc, _stream())
        _lib.check(st, 'dl_stat_pool')
    else:
        st = _lib.lib().dl_attn_stat_pool(_ptr(x), _ptr(logits), B, T, Cc, ldx, _ptr(lengths), _ptr(of), _ptr(ob),
                                          2 * Cc, _stream())
        _lib.check(st, 'dl_attn_stat_pool')
    return of, ob


def attn_logits(h_f32, v, k):
    _need_cuda(h_f32, v)
    rows, Hd = h_f32.shape
    e = torch.empty((rows,), device=h_f32.device, dtype=torch.float32)
    st = _lib.lib().dl_attn_logits(_ptr(h_f32), rows, Hd, Hd, _ptr(v), float(k), _ptr(e), _stream())
    _lib.check(st, 'dl_attn_logits')
    return e


# ------------------------------------------------------------------ K8
def znorm_concat(a, v, biased=False, video_first=False, l2norm=False, want_bf16=False, out=None):
    """out: optional caller-owned contiguous (B, Da+Dv) f32 rows -- e.g. this batch's slice of the all-gather
    table (K8 writes straight into the collective's buffer, no copy before the all_gather)."""
    _need_cuda(a, v, out)
    a = a.contiguous().float()
    v = v.contiguous().float()
    B, Da = a.shape
    Dv = v.shape[1]
    if out is None:
        out = torch.empty((B, Da + Dv), device=a.device, dtype=torch.float32)
    else:
        assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, Da + Dv)
    ob = torch.empty((B, Da + Dv), device=a.device, dtype=torch.bfloat16) if want_bf16 else None
    st = _lib.lib().dl_znorm_concat(_ptr(a), Da, _ptr(v), Dv, B, int(biased), int(video_first), int(l2norm),
                                    _ptr(out), _ptr(ob), _stream())
    _lib.check(st, 'dl_znorm_concat')
    return (out, ob) if want_bf16 else out


def lowfer(e1, e2):
    _need_cuda(e1, e2)
    e1 = e1.contiguous().float()
    e2 = e2.contiguous().float()
    B, D = e1.shape
    out = torch.empty((B, 3 * D), device=e1.device, dtype=torch.float32)
    _lib.check(_lib.lib().dl_lowfer(_ptr(e1), _ptr(e2), B, D, _ptr(out), _stream()), 'dl_lowfer')
    return out


def l2_normalize(x, want_bf16=False):
    _need_cuda(x)
    x = x.contiguous().float()
    B, D = x.shape
    out = torch.empty_like(x)
    ob = torch.empty((B, D), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _lib.check(_lib.lib().dl_l2_normalize(_ptr(x), B, D, _ptr(out), _ptr(ob), _stream()), 'dl_l2_normalize')
    return (out, ob) if want_bf16 else out


def affine_act(x, scale=None, shift=None, slope=1.0, ld=None, want_bf16=True, want_f32=False):
    """lrelu(x*scale+shift) on (rows,C) f32 -> (bf16 (rows,ld) | None, f32 (rows,C) | None)."""
    _need_cuda(x, scale, shift)
    x = x.contiguous().float()
    rows, Cc = x.shape
    ld = ld or Cc
    y = torch.empty((rows, ld), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    yf = torch.empty((rows, Cc), device=x.device, dtype=torch.float32) if want_f32 else None
    st = _lib.lib().dl_affine_act(_ptr(x), rows, Cc, _ptr(scale), _ptr(shift), float(slope), _ptr(y), ld, _ptr(yf),
                                  _stream())
    _lib.check(st, 'dl_affine_act')
    return y, yf


# ------------------------------------------------------------------ K9
def cosine_score_trials(emb, enrol, test):
    _need_cuda(emb, enrol, test)
    emb = emb.contiguous().float()
    assert enrol.dtype == torch.int32 and test.dtype == torch.int32
    n = enrol.numel()
    scores = torch.empty((n,), device=emb.device, dtype=torch.float32)
    st = _lib.lib().dl_cosine_score_trials(_ptr(emb), emb.shape[0], emb.shape[1], _ptr(enrol), _ptr(test), n,
                                           _ptr(scores), _stream())
    _lib.check(st, 'dl_cosine_score_trials')
    return scores


def score_fusion_trials(emb_a, emb_v, enrol, test):
    _need_cuda(emb_a, emb_v, enrol, test)
    emb_a = emb_a.contiguous().float()
    emb_v = emb_v.contiguous().float()
    n = enrol.numel()
    scores = torch.empty((n,), device=emb_a.device, dtype=torch.float32)
    st = _lib.lib().dl_score_fusion_trials(_ptr(emb_a), emb_a.shape[1], _ptr(emb_v), emb_v.shape[1], emb_a.shape[0],
                                           _ptr(enrol), _ptr(test), n, _ptr(scores), _stream())
    _lib.check(st, 'dl_score_fusion_trials')
    return scores


def gather_scores(S, rows, cols):
    _need_cuda(S, rows, cols)
    n = rows.numel()
    out = torch.empty((n,), device=S.device, dtype=torch.float32)
    st = _lib.lib().dl_gather_scores(_ptr(S), S.stride(0), _ptr(rows), _ptr(cols), n, _ptr(out), _stream())
    _lib.check(st, 'dl_gather_scores')
    return out
