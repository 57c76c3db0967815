import sys; sys.path.insert(0, '/root/repo')
import torch, bench
from deeplip_b200 import ops
from deeplip_b200.pipeline import build_models
B = 64
audio, video = build_models('cuda', seed=1)
raw, wav = bench.synth_batch(B, seed=1)
raw = torch.from_numpy(raw).cuda()
def mx(a, b): return float((a.float() - b.float()).abs().max())
tr = video.trunk
def run_pass(keep):
    pk = video._packed()
    bufs = tr.stacked_buffers(B * 75, 22, 22, raw.device, 5)
    ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], out=bufs[-1])
    keep.append(bufs[-1].clone())
    x = bufs[-1]
    for i, blk in enumerate(tr.layer1):
        x = blk.forward_stacked(x, 22, bufs[2 * i], bufs[2 * i + 1])
        keep.append(x.clone())
    x = tr.layer2[0].forward_nhwc(x, H=22); keep.append(x)
    for blk in list(tr.layer2)[1:] + list(tr.layer3) + list(tr.layer4):
        x = blk.forward_nhwc(x); keep.append(x)
    return x
for mode in ('nosync', 'sync'):
    passes = []
    for i in range(3):
        k = []; run_pass(k); passes.append(k)
        if mode == 'sync': torch.cuda.synchronize()
    torch.cuda.synchronize()
    names = ['stem', 'l1.0', 'l1.1', 'l2.0', 'l2.1', 'l3.0', 'l3.1', 'l4.0', 'l4.1']
    print(mode, [(n, mx(a, b), mx(b, c)) for n, a, b, c in zip(names, *passes)])
