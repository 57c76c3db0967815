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
cases = []
for slots in (6,):
    cases += [(1024, 'no pooling (BN+ring only)'), (2048, 'no BN/ring (pooling only)'), (0 | 48 | 8, 'MMA only 1/4'), (slots << 8, '%d slots full' % slots), (0 | 48, '%d slots, MMA only' % slots), (0 | 32, '%d slots, epilogue idle' % slots), (0 | 16, '%d slots, builders idle' % slots)]
for dbg, what in cases:
    _lib.set_option('dbg', dbg)
    t = timeit(lambda: ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a']), n=(1 if dbg & 512 else 10)); torch.cuda.synchronize()
    print('stem %-24s %7.1f us (incl. prepass)' % (what, t), flush=True)
_lib.set_option('dbg', 0)
