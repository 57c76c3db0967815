import sys; sys.path.insert(0, '/root/repo')
import torch, bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import build_models
B = 64
audio, video = build_models('cuda', seed=1)
raw, wav = bench.synth_batch(B, seed=1)
raw = torch.from_numpy(raw).cuda()
def mx(a, b): return float((a.float() - b.float()).abs().max())
pk = video._packed()
N, Hp = B * 75, 22
x = torch.zeros((N, Hp + 1, Hp, 64), device='cuda', dtype=torch.bfloat16)
ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], crop=(88, 88), out=x)
blk = video.trunk.layer1[0]
p = blk._packed()
mid = torch.zeros_like(x); out = torch.zeros_like(x)
r = []
for i in range(4):
    ops.conv3x3_halo(x, p['w1'], p['s1'], p['h1'], p['a1'], Hp, out=mid)
    r.append(mid.clone())
torch.cuda.synchronize()
print('conv1 same out buffer, back-to-back:', [mx(r[0], r[i]) for i in range(1, 4)])
r = []
for i in range(4):
    ops.conv3x3_halo(mid, p['w2'], p['s2'], p['h2'], p['a2'], Hp, out=out, residual=x)
    r.append(out.clone())
torch.cuda.synchronize()
print('conv2+res same out buffer, back-to-back:', [mx(r[0], r[i]) for i in range(1, 4)])
r = []
for i in range(4):
    ops.conv3x3_halo(mid, p['w2'], p['s2'], p['h2'], p['a2'], Hp, out=out, residual=x)
    torch.cuda.synchronize()
    r.append(out.clone())
print('conv2+res synced:', [mx(r[0], r[i]) for i in range(1, 4)])
# igemm reference of the same conv for ground truth (dense layout)
xd = x[:, :Hp].contiguous(); md = mid[:, :Hp].contiguous()
ref, _ = ops.conv_igemm(md, p['w2'], 64, 64, 3, 3, (1, 1), (1, 1), (1, 1), p['s2'], p['h2'], p['a2'], residual=xd)
print('halo vs igemm (each run):', [mx(ri[:, :Hp], ref) for ri in r])
# where do the unsynced runs differ from the reference?
r = []
ops.conv3x3_halo(x, p['w1'], p['s1'], p['h1'], p['a1'], Hp, out=mid)
for i in range(3):
    ops.conv3x3_halo(mid, p['w2'], p['s2'], p['h2'], p['a2'], Hp, out=out, residual=x)
    r.append(out.clone())
torch.cuda.synchronize()
for i, ri in enumerate(r):
    d = (ri[:, :Hp].float() - ref.float()).abs()
    bad = (d > 0).nonzero()
    print('run', i, 'n bad', bad.shape[0], 'max', float(d.max()))
    if bad.shape[0]:
        R = bad[:, 0] * 23 + bad[:, 1]           # stacked row
        tiles = torch.unique(R // 16)
        print('   images', torch.unique(bad[:, 0])[:10].tolist(), 'rows', torch.unique(bad[:, 1]).tolist()[:23],
              'cols', torch.unique(bad[:, 2]).tolist(), 'n row-tiles', tiles.numel(), 'first tiles', tiles[:12].tolist(),
              'chan range', int(bad[:, 3].min()), int(bad[:, 3].max()))
