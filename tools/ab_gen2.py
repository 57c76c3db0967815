"""A/B of the kernel generations selectable with dl_set_option (front end 1|2, stem pre-pass 1|2) plus the small
HBM-bound kernels timed alone: CUDA events around each call, 256 MiB L2 flush in between, median of 10."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from deeplip_b200 import _lib, ops
from deeplip_b200.pipeline import build_models

dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t_ms(fn, reps=10):
    fn(); fn()
    ev = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


raw, wav = bench.synth_batch(64, seed=1)
raw, wav = torch.from_numpy(raw).to(dev), torch.from_numpy(wav).to(dev)
big = wav.repeat(16, 1)
for gen in (1, 2):
    _lib.set_option('frontend', gen)
    for name, w in (('B=64', wav), ('B=1024', big)):
        for ft, nf in (('mfcc', 24), ('logfbank', 60)):
            ms = t_ms(lambda: ops.frontend_features(w, ft, nf, True))
            by = w.shape[0] * (48000 * 4 + 299 * nf * 4)
            print('frontend gen%d %-7s %-8s %8.1f us  %7.1f GB/s (algorithmic)' % (gen, name, ft, ms * 1e3, by / ms / 1e6))
_lib.set_option('frontend', 2)
ms = t_ms(lambda: ops.frontend_features(wav, 'stft', 257, True))
print('frontend gen2 B=64    stft     %8.1f us' % (ms * 1e3))
a1 = ops.frontend_features(wav, 'mfcc', 24, True)
_lib.set_option('frontend', 1)
a0 = ops.frontend_features(wav, 'mfcc', 24, True)
_lib.set_option('frontend', 2)
print('gen1 vs gen2 max abs diff f32 %.3e  bf16 %.3e' % (float((a1[0] - a0[0]).abs().max()),
                                                          float((a1[1].float() - a0[1].float()).abs().max())))

audio, video = build_models(dev, seed=1)
pk = video._packed()
for gen in (1, 2):
    _lib.set_option('prepass', gen)
    ms = t_ms(lambda: ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], crop=(88, 88)))
    print('stem (pre-pass gen%d + conv3d) B=64 %8.1f us' % (gen, ms * 1e3))
_lib.set_option('prepass', 2)

xs = torch.randn(64, 277, 1504, device=dev).to(torch.bfloat16)
ms = t_ms(lambda: ops.stat_pool(xs, 1500))
print('stat_pool B=64 %8.1f us  %7.1f GB/s' % (ms * 1e3, 64 * 1504 * 277 * 2 / ms / 1e6))
xf = torch.randn(64 * 75, 3, 3, 512, device=dev).to(torch.bfloat16)
try:
    ms = t_ms(lambda: ops.frame_pool_temporal_mean(xf, 64, 75))
    print('frame_pool B=64 %8.1f us  %7.1f GB/s' % (ms * 1e3, xf.numel() * 2 / ms / 1e6))
except Exception as e:
    print('frame_pool: skipped', repr(e)[:120])
ms = t_ms(lambda: audio.embed_ntc(a1[1]))
print('E-TDNN + pooling + heads B=64 %8.1f us' % (ms * 1e3))
ms = t_ms(lambda: video.utterance_embedding(raw))
print('video branch B=64 %8.1f us' % (ms * 1e3))

# ---- heads: linear_small_kernel vs the tensor-core igemm; stat pool loads in flight; host -> device bandwidth
pooled = torch.randn(64, 3008, device=dev).to(torch.bfloat16)
pk_a = audio._packed()
E = audio.embedding_dim
for sl in (0, 1):
    _lib.set_option('small_linear', sl)
    ms1 = t_ms(lambda: ops.conv_igemm(pooled.view(64, 1, 1, -1), pk_a['w1'], pooled.shape[1], E, scale=pk_a['s1'],
                                      shift=pk_a['h1'], slope=pk_a['lrelu'], want_f32=True, scale2=pk_a['one'],
                                      shift2=pk_a['b1']))
    h = torch.randn(64, 1, 1, E, device=dev).to(torch.bfloat16)
    ms2 = t_ms(lambda: ops.conv_igemm(h, pk_a['w2'], E, E, want_bf16=False, want_f32=True, scale2=pk_a['one'],
                                      shift2=pk_a['b2']))
    print('heads small_linear=%d: fc1 %6.1f us  fc2 %6.1f us' % (sl, ms1 * 1e3, ms2 * 1e3))
_lib.set_option('small_linear', 1)
for mlp in (4, 8):
    _lib.set_option('statpool_mlp', mlp)
    ms = t_ms(lambda: ops.stat_pool(xs, 1500))
    print('stat_pool loads in flight %d: %6.1f us' % (mlp, ms * 1e3))
_lib.set_option('statpool_mlp', 4)
for slab in (256, 128):
    _lib.set_option('statpool_slab', slab)
    ms = t_ms(lambda: ops.stat_pool(xs, 1500))
    print('stat_pool slab %d: %6.1f us  %7.1f GB/s' % (slab, ms * 1e3, 64 * 1504 * 277 * 2 / ms / 1e6))
_lib.set_option('statpool_slab', 256)
hraw = torch.empty((64, 75, 96, 96), dtype=torch.uint8).pin_memory()
draw = torch.empty_like(hraw, device=dev)
for _ in range(3):
    draw.copy_(hraw, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    draw.copy_(hraw, non_blocking=True)
b.record(); torch.cuda.synchronize()
print('H2D pinned 44 MB: %.2f ms each = %.1f GB/s' % (a.elapsed_time(b) / 10, hraw.numel() * 10 / a.elapsed_time(b) / 1e6))

from deeplip_b200.video_models import resnet as R
for fuse in (False, True):
    R.FUSE_L2_ENTRY = fuse
    ms = t_ms(lambda: video.utterance_embedding(raw))
    print('video branch B=64, layer2 entry fused=%d: %8.1f us' % (fuse, ms * 1e3))
R.FUSE_L2_ENTRY = True
