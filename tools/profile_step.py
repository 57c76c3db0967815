"""Minimal driver for ncu captures: N iterations of the video branch (stem + trunk) and the audio branch at B=64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from deeplip_b200.pipeline import AVExtractor, build_models
B = int(os.environ.get('B', '64'))
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
from deeplip_b200 import _lib
for k in ('pair', 'pair_resident'):
    if os.environ.get('DL_OPT_' + k.upper()) is not None:
        _lib.set_option(k, int(os.environ['DL_OPT_' + k.upper()]))
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(B, seed=1)
raw, wav = torch.from_numpy(raw).cuda(), torch.from_numpy(wav).cuda()
for _ in range(iters):
    ex.extract(wav, raw)
torch.cuda.synchronize()
print('done')
