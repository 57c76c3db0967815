"""Timing ablations of the second-generation stem at B = 64 (dbg bits of stem2_conv3d.cuh).  python tools/stem2_abl.py [ncu]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, _lib, synth
from deeplip_b200.pipeline import build_models
from lin_bench import timeit
_, video = build_models()
pk = video._packed()
B, T = 64, 75
x = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=96, W=96, seed=3)).cuda()
f = lambda: ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a'])
if len(sys.argv) > 1 and sys.argv[1] == 'ncu':
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    sys.exit(0)
cases = [(0, 'full')] + [(int(a, 0), 'dbg ' + a) for a in sys.argv[1:]]
outs = {}
for rnd in range(2):
    for dbg, what in cases:
        _lib.set_option('dbg', dbg)
        t = timeit(f, n=10); torch.cuda.synchronize()
        outs[dbg] = f().clone()
        print('stem gen2 %-28s %7.1f us (incl. prepass)' % (what, t), flush=True)
_lib.set_option('dbg', 0)
for d, o in outs.items():
    if d in (2048, 4096):
        print('dbg %d output bit-identical to the default:' % d, torch.equal(o, outs[0]))
