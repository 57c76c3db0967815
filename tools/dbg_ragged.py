import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from deeplip_b200 import _lib, synth, ops
from deeplip_b200.pipeline import AVExtractor, build_models, HostPipeline
from oracle import models_ref
DEV = 'cuda'
audio, video = build_models(DEV, seed=1)
ex = AVExtractor(audio, video)
B, T, nsamp = 4, 8, 24000
spk = [1, 1, 2, 3]
wav = synth.speech_like_audio(spk, nsamp=nsamp, seed=1)
raw = synth.lip_crops_u8(spk, T=T, seed=1)
wl = torch.tensor([nsamp, nsamp - 5000] + [nsamp] * (B - 2), dtype=torch.int32, device=DEV)
vl = torch.tensor([T, T - 3] + [T] * (B - 2), dtype=torch.int32, device=DEV)
wav2 = wav.copy(); wav2[1, nsamp - 5000:] = 0
raw_f = torch.stack([models_ref.video_preprocess(torch.from_numpy(r)) for r in raw]); raw_f[1, T - 3:] = 0
for sl in (0, 1):
    for mlp in (4, 8):
        _lib.set_option('small_linear', sl); _lib.set_option('statpool_mlp', mlp)
        for fus in ('audio', 'video', 'concat'):
            e = AVExtractor(audio, video, fusion=fus)
            rag = e.extract(torch.from_numpy(wav2).to(DEV), raw_f.to(DEV), wl, vl)
            alone = e.extract(torch.from_numpy(wav2[1:2, :nsamp - 5000]).to(DEV), raw_f[1:2, :T - 3].contiguous().to(DEV))
            torch.cuda.synchronize()
            print('small_linear', sl, 'mlp', mlp, fus, 'ragged_abs', float((rag[1] - alone[0]).abs().max()))
_lib.set_option('small_linear', 1); _lib.set_option('statpool_mlp', 4)

# ---- e2e: where do the 12 ms per step go?
import bench
hostb = []
for r in range(4):
    rw, wv = bench.synth_batch(64, seed=r + 1)
    hostb.append((torch.from_numpy(wv).pin_memory(), torch.from_numpy(rw).pin_memory()))
hp = HostPipeline(ex, torch.device('cuda', 0))
for n in (2, 10, 10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hp.run([hostb[i % 4] for i in range(n)])
    torch.cuda.synchronize(); print('hp.run %d batches: %.2f ms/step' % (n, (time.perf_counter() - t0) * 1e3 / n))
dv = [(a.cuda(), b.cuda()) for a, b in hostb]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(10):
    ex.extract(*dv[i % 4])
torch.cuda.synchronize(); print('device-resident extract: %.2f ms/step' % ((time.perf_counter() - t0) * 100))
t0 = time.perf_counter()
for i in range(10):
    w, v = hostb[i % 4]
    ex.extract(w.to(DEV, non_blocking=True), v.to(DEV, non_blocking=True))
torch.cuda.synchronize(); print('naive h2d + extract: %.2f ms/step' % ((time.perf_counter() - t0) * 100))
t0 = time.perf_counter()
for i in range(10):
    w, v = hostb[i % 4]
    w.to(DEV, non_blocking=True); v.to(DEV, non_blocking=True)
torch.cuda.synchronize(); print('h2d only: %.2f ms/step' % ((time.perf_counter() - t0) * 100))
