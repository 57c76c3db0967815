"""A/B the tuning switches inside ONE process (same box, same clocks): trunk time per option setting."""
import sys; sys.path.insert(0, '/root/repo')
import statistics, torch, bench
from deeplip_b200 import _lib, ops
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
    for pair, res in ((1, 1), (1, 0), (0, 0)):
        _lib.set_option('pair', pair); _lib.set_option('pair_resident', res)
        print('round %d pair=%d resident=%d  trunk %.3f ms  audio %.3f ms  step %.3f ms' % (
            rnd, pair, res, t(lambda: video.trunk.forward_nhwc(buf, stacked_H=22)), t(lambda: ex.audio_embedding(wav)),
            t(lambda: ex.extract(wav, raw))))
