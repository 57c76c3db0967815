import sys; sys.path.insert(0, '/root/repo')
import torch, bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import build_models
B = 64
audio, video = build_models('cuda', seed=1)
raw, wav = bench.synth_batch(B, seed=1)
raw = torch.from_numpy(raw).cuda()
pk = video._packed()
N, Hp = B * 75, 22
def mx(a, b): return float((a.float() - b.float()).abs().max())
outs = []
for i in range(3):
    buf = torch.zeros((N, Hp + 1, Hp, 64), device='cuda', dtype=torch.bfloat16)
    ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], crop=(88, 88), out=buf)
    outs.append(buf)
torch.cuda.synchronize()
print('stem->stacked repeat diffs', mx(outs[0], outs[1]), mx(outs[0], outs[2]), 'pad', float(outs[0][:, 22:].float().abs().max()))
dense = ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], crop=(88, 88))
print('stacked vs dense', mx(outs[0][:, :22], dense))
x = outs[0]
blk = video.trunk.layer1[0]
res = []
for i in range(3):
    mid = torch.zeros_like(x); out = torch.zeros_like(x)
    blk.forward_stacked(x, Hp, mid, out)
    res.append((mid, out))
torch.cuda.synchronize()
print('halo conv1 repeat', mx(res[0][0], res[1][0]), mx(res[0][0], res[2][0]))
print('halo conv2(res) repeat', mx(res[0][1], res[1][1]), mx(res[0][1], res[2][1]))
t = [video.trunk.forward_nhwc(x, stacked_H=Hp).clone() for _ in range(3)]
torch.cuda.synchronize()
print('trunk(stacked) repeat', mx(t[0], t[1]), mx(t[0], t[2]), mx(t[1], t[2]))
m = [video.trunk_maps(raw).clone() for _ in range(3)]
torch.cuda.synchronize()
print('trunk_maps repeat', mx(m[0], m[1]), mx(m[1], m[2]), 'vs trunk(stacked)', mx(m[0], t[0]), mx(m[2], t[0]))
