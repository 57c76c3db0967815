"""Run every stage of the video/audio path repeatedly on the same input and report bitwise differences."""
import sys; sys.path.insert(0, '/root/repo')
import torch, bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import AVExtractor, build_models
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(B, seed=1)
raw, wav = torch.from_numpy(raw).cuda(), torch.from_numpy(wav).cuda()
pk = video._packed()
def stem():
    return ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], crop=(88, 88)).clone()
s0 = stem()
for i in range(3):
    s1 = stem(); torch.cuda.synchronize()
    d = (s1.float() - s0.float()).abs()
    print('stem run', i, 'max diff', float(d.max()), 'n diff', int((d > 0).sum()))
def trunk(x):
    return video.trunk.forward_nhwc(x.clone()).clone()
t0 = trunk(s0)
for i in range(3):
    t1 = trunk(s0); torch.cuda.synchronize()
    d = (t1.float() - t0.float()).abs()
    print('trunk run', i, 'max diff', float(d.max()), 'n diff', int((d > 0).sum()))
# layer by layer on the dense path pieces
x = s0
for li, layer in enumerate([video.trunk.layer1, video.trunk.layer2, video.trunk.layer3, video.trunk.layer4]):
    for bi, blk in enumerate(layer):
        import os
        y0 = blk.forward_nhwc(x).clone()
        y1 = blk.forward_nhwc(x).clone(); torch.cuda.synchronize()
        d = (y1.float() - y0.float()).abs()
        print('layer%d.%d igemm max diff %g n %d' % (li + 1, bi, float(d.max()), int((d > 0).sum())))
        x = y0
a0 = ex.audio_embedding(wav).clone()
for i in range(2):
    a1 = ex.audio_embedding(wav).clone(); torch.cuda.synchronize()
    print('audio run', i, 'max diff', float((a1 - a0).abs().max()))
e0 = ex.extract(wav, raw).clone()
for i in range(3):
    e1 = ex.extract(wav, raw).clone(); torch.cuda.synchronize()
    print('extract run', i, 'max diff', float((e1 - e0).abs().max()))
