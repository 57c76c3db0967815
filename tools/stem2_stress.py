"""Bitwise stress of the second-generation stem: N back-to-back launches at B = 64 into rotating buffers, every result
compared with the first bit for bit (a missing proxy fence or a barrier-phase slip shows up as a few corrupted tiles in
some launches, invisible to tolerance parity).   python tools/stem2_stress.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, synth
from deeplip_b200.pipeline import build_models
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
_, video = build_models()
pk = video._packed()
x = torch.from_numpy(synth.lip_crops_u8([1] * 64, T=75, H=96, W=96, seed=3)).cuda()
f = lambda out=None: ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a'], out=out)
ref = f().clone()
torch.cuda.synchronize()
bufs = [torch.empty_like(ref) for _ in range(4)]
bad = 0
for i in range(0, N, 4):
    for b in bufs:
        b.fill_(7.0)
    for b in bufs:
        f(out=b)
    torch.cuda.synchronize()
    bad += sum(int(not torch.equal(b, ref)) for b in bufs)
print('stem2: %d back-to-back launches, results differing from the first: %d' % (N, bad))
assert bad == 0
print('ok')
