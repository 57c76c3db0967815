"""Second-generation stem (stem2_conv3d.cuh) against the first: parity cases, bit comparison, timing at B = 64.
    python tools/stem2_try.py"""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from deeplip_b200 import ops, _lib, synth
from deeplip_b200.pipeline import build_models
from lin_bench import timeit
import gpu_checks as G

_, video = build_models()
pk = video._packed()


def run(x, opt, **kw):
    _lib.set_option('stem', opt)
    try:
        y = ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a'], **kw)
        torch.cuda.synchronize()
    finally:
        _lib.set_option('stem', 2)
    return y


for B, T in ((1, 1), (2, 6), (2, 5), (3, 75)):
    x = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=96, W=96, seed=B + T)).cuda()
    try:
        y1, y2 = run(x, 1), run(x, 2)
        d = (y1.float() - y2.float()).abs()
        nb = int((y1.view(torch.int16) != y2.view(torch.int16)).sum())
        print('B=%d T=%d: gen1 vs gen2 max abs %.4g, differing bf16 values %d of %d, gen2 finite %s' %
              (B, T, float(d.max()), nb, y1.numel(), bool(torch.isfinite(y2.float()).all())), flush=True)
        if nb:
            idx = torch.nonzero(y1.view(torch.int16) != y2.view(torch.int16))
            print('   first differing (frame, py, px, ch):', idx[:6].tolist(), ' last:', idx[-3:].tolist(), flush=True)
            fr = torch.unique(idx[:, 0]).tolist(); print('   frames', fr[:20], 'py', torch.unique(idx[:, 1]).tolist(), 'px', torch.unique(idx[:, 2]).tolist()[:30], 'ch', torch.unique(idx[:, 3]).tolist()[:70], flush=True)
    except Exception:
        traceback.print_exc()
for kw in (dict(B=2, T=6), dict(B=2, T=5, u8=True), dict(B=1, T=3, H=32, W=32)):
    try:
        print('stem_case', kw, G.stem_case(**kw), flush=True)
    except Exception:
        traceback.print_exc()
# ragged lengths + stacked-rows output
try:
    x = torch.from_numpy(synth.lip_crops_u8([1] * 3, T=9, H=96, W=96, seed=5)).cuda()
    ln = torch.tensor([9, 4, 7], dtype=torch.int32, device='cuda')
    outs = []
    for opt in (1, 2):
        out = torch.zeros(27, 23, 22, 64, device='cuda', dtype=torch.bfloat16)
        run(x, opt, out=out, lengths=ln)
        outs.append(out)
    print('ragged + stacked rows: max abs diff', float((outs[0].float() - outs[1].float()).abs().max()),
          'pad row zero', bool((outs[1][:, 22] == 0).all()), flush=True)
except Exception:
    traceback.print_exc()

B, T = 64, 75
x = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=96, W=96, seed=3)).cuda()
for rnd in range(2):
    for opt in (1, 2):
        _lib.set_option('stem', opt)
        t = timeit(lambda: ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a']), n=10); torch.cuda.synchronize()
        print('stem gen%d full %7.1f us (incl. prepass)' % (opt, t), flush=True)
_lib.set_option('stem', 2)
for dbg, what in ((32 | 2, 'no staging writes, no bulk stores'), (64, 'no epilogue arithmetic'), (8, '1/4 of the MMAs'), (16, 'builders skip the strip reads'),
                  (128, 'no strip TMA loads'), (128 | 16, 'no strip loads, no strip reads'), (128 | 64, 'no strip loads, no epilogue arithmetic'),
                  (128 | 16 | 64 | 32 | 2, 'MMAs + barriers only'), (128 | 8 | 16 | 64 | 32 | 2, 'pipeline only')):
    _lib.set_option('dbg', dbg)
    t = timeit(lambda: ops.stem_conv3d(x, pk['w'], pk['s'], pk['h'], pk['a']), n=10); torch.cuda.synchronize()
    print('stem gen2 %-36s %7.1f us (incl. prepass)' % (what, t), flush=True)
_lib.set_option('dbg', 0)
