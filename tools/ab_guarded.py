"""A/B inside ONE process: guarded layer2 on / off, stem, trunk, audio and whole-step times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import statistics, torch, bench
from deeplip_b200 import _lib, ops
from deeplip_b200.video_models import resnet as R
from deeplip_b200.pipeline import AVExtractor, build_models
B = 64
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(B, seed=1)
raw, wav = torch.from_numpy(raw).cuda(), torch.from_numpy(wav).cuda()
pk = video._packed()
buf = video.trunk.stacked_buffers(B * 75, 22, 22, raw.device, 5)[-1]
ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], out=buf)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def t(fn, n=15):
    for _ in range(3): fn()
    evs = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)
for rnd in range(2):
    for g in (1, 0):
        R.USE_GUARDED = bool(g)
        print('round %d guarded=%d  stem %.3f  trunk %.3f ms  audio %.3f ms  step %.3f ms' % (
            rnd, g, t(lambda: ops.stem_conv3d(raw, pk['w'], pk['s'], pk['h'], pk['a'], out=buf)),
            t(lambda: video.trunk.forward_nhwc(buf, stacked_H=22)), t(lambda: ex.audio_embedding(wav)),
            t(lambda: ex.extract(wav, raw))), flush=True)
