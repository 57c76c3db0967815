"""Does the video trunk run faster on utterance chunks whose activations stay L2-resident?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import synth
from deeplip_b200.pipeline import build_models

_, video = build_models()
B, T = 64, 29
x = torch.from_numpy(synth.lip_crops_u8([1] * B, T=T, H=96, W=96, seed=3)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def run(chunk):
    outs = [video.utterance_embedding(x[i:i + chunk]) for i in range(0, B, chunk)]
    return torch.cat(outs)


ref = run(B)
for chunk in (64, 32, 16, 8, 4):
    out = run(chunk)
    torch.cuda.synchronize()
    err = float((out - ref).abs().max())
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(chunk)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print('chunk %2d utt: video trunk %.3f ms (min %.3f)  max|diff| vs full batch %.2e' % (chunk, ts[len(ts) // 2], ts[0], err),
          flush=True)
