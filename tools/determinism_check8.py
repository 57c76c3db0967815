import sys, os; sys.path.insert(0, '/root/repo')
import torch, bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import build_models
B = 64
audio, video = build_models('cuda', seed=1)
raw, wav = bench.synth_batch(B, seed=1)
raw = torch.from_numpy(raw).cuda()
pk = video._packed()
N, Hp = B * 75, 22
x = torch.zeros((N, Hp + 1, Hp, 64), device='cuda', dtype=torch.bfloat16)
ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], crop=(88, 88), out=x)
p = video.trunk.layer1[0]._packed()
xd = x[:, :Hp].contiguous()
midd, _ = ops.conv_igemm(xd, p['w1'], 64, 64, 3, 3, (1, 1), (1, 1), (1, 1), p['s1'], p['h1'], p['a1'])
ref, _ = ops.conv_igemm(midd, p['w2'], 64, 64, 3, 3, (1, 1), (1, 1), (1, 1), p['s2'], p['h2'], p['a2'], residual=xd)
mid = torch.zeros_like(x); out = torch.zeros_like(x)
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
f1 = 0
for it in range(n):
    ops.conv3x3_halo(x, p['w1'], p['s1'], p['h1'], p['a1'], Hp, out=mid)
    torch.cuda.synchronize()
    if int(((mid[:, :Hp].float() - midd.float()).abs() > 0).sum()): f1 += 1
f2 = 0
for it in range(n):
    ops.conv3x3_halo(mid, p['w2'], p['s2'], p['h2'], p['a2'], Hp, out=out, residual=x)
    torch.cuda.synchronize()
    if int(((out[:, :Hp].float() - ref.float()).abs() > 0).sum()): f2 += 1
print('DL_HALO_LDG_RES=%s  conv1 (no residual) failures %d/%d   conv2 (+residual) failures %d/%d' % (os.environ.get('DL_HALO_LDG_RES'), f1, n, f2, n))
