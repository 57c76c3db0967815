import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import ops, packing, _lib, synth
from deeplip_b200.pipeline import build_models
from lin_bench import timeit
_, video = build_models()
pk = video._packed()
B, T = 64, 75
x = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=96, W=96, seed=3)).cuda()
cases = [(0, 'full'), (48, 'MMA only'), (48 | 8, 'MMA only 1/4'), (48 | 16384, 'MMA only, no strip loads'), (48 | 8 | 16384, 'MMA 1/4, no strip loads'),
         (32 | 16384, 'epilogue idle, no strip loads'), (16384, 'full, no strip loads (garbage)')]
for dbg, what in cases:
    _lib.set_option('dbg', dbg)
    t = timeit(lambda: ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a']), n=(1 if dbg & 512 else 10)); torch.cuda.synchronize()
    print('stem %-24s %7.1f us (incl. prepass)' % (what, t), flush=True)
_lib.set_option('dbg', 0)
