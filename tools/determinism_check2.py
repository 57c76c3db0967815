import sys; sys.path.insert(0, '/root/repo')
import torch, bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import AVExtractor, build_models
B = 64
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(B, seed=1)
raw, wav = torch.from_numpy(raw).cuda(), torch.from_numpy(wav).cuda()
ems = []
for i in range(4):
    y = video.trunk_maps(raw).clone()
    em = video.utterance_embedding(raw).clone()
    torch.cuda.synchronize()
    ems.append((y, em))
    if i:
        print('call', i, 'maps diff vs call0', float((y.float() - ems[0][0].float()).abs().max()),
              'em diff', float((em - ems[0][1]).abs().max()),
              'maps diff vs prev', float((y.float() - ems[i-1][0].float()).abs().max()))
bufs = video.trunk._stk
print('pad rows abs max per buffer', [float(b[:, 22:].float().abs().max()) for b in bufs])
xv = [ex.audio_embedding(wav).clone() for _ in range(3)]
print('audio diffs', float((xv[1]-xv[0]).abs().max()), float((xv[2]-xv[0]).abs().max()))
es = [ex.extract(wav, raw).clone() for _ in range(3)]
print('extract diffs', float((es[1]-es[0]).abs().max()), float((es[2]-es[1]).abs().max()),
      'audio half', float((es[1][:, :512]-es[0][:, :512]).abs().max()), 'video half', float((es[1][:, 512:]-es[0][:, 512:]).abs().max()))
