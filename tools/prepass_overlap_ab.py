"""Stem pre-pass on a side stream under the audio branch (AVExtractor.overlap_prepass) against the plain order:
interleaved A/B of the whole step at B = 64, rotating inputs.   python tools/prepass_overlap_ab.py"""
import os, sys, statistics, time
BURST = len(sys.argv) > 1 and sys.argv[1] == 'burst'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from deeplip_b200.pipeline import AVExtractor, build_models
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
batches = []
for r in range(4):
    raw, wav = bench.synth_batch(64, seed=r + 1)
    batches.append((torch.from_numpy(wav).cuda(), torch.from_numpy(raw).cuda()))
def run(flag, n=20):
    ex.overlap_prepass = flag
    torch.cuda.synchronize()
    if BURST:
        time.sleep(2.0)          # bench.py's protocol: every timed region starts from an idle GPU
    for i in range(3):
        ex.extract(*batches[i % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        ex.extract(*batches[i % 4])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ex.overlap_prepass = False
ref = ex.extract(*batches[0]).clone()
ex.overlap_prepass = True
got = ex.extract(*batches[0]).clone()
torch.cuda.synchronize()
print('bit-identical:', torch.equal(ref, got))
res = {False: [], True: []}
for rnd in range(6):
    for flag in (False, True):
        res[flag].append(run(flag))
for flag in (False, True):
    print('overlap_prepass=%s: ms/step per round %s  median %.3f' % (flag, ['%.3f' % v for v in res[flag]], statistics.median(res[flag])))
