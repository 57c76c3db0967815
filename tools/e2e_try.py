"""Where does the end-to-end step (pinned host -> HostPipeline -> pinned host) lose against the device-resident step?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from deeplip_b200.pipeline import AVExtractor, HostPipeline, build_models
dev = torch.device('cuda', 0)
audio, video = build_models(dev, seed=1)
ex = AVExtractor(audio, video)
hostb = []
for r in range(4):
    rw, wv = bench.synth_batch(64, seed=r + 1)
    hostb.append((torch.from_numpy(wv).pin_memory(), torch.from_numpy(rw).pin_memory()))
dv = [(a.to(dev), b.to(dev)) for a, b in hostb]


def ev_time(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)


def dev_loop(n):
    for i in range(n):
        ex.extract(*dv[i % 4])


for _ in range(2):
    dev_loop(10)
for n in (10, 40):
    print('device-resident, no flush, %2d steps: %.3f ms/step' % (n, ev_time(lambda: dev_loop(n)) / n))
for slots in (2, 3):
    hp = HostPipeline(ex, dev, slots=slots)
    hp.run([hostb[i % 4] for i in range(10)])
    for n in (10, 40):
        hp.run([hostb[i % 4] for i in range(n)])
        print('HostPipeline slots=%d, %2d steps: %.3f ms/step' % (slots, n, ev_time(lambda: hp.run([hostb[i % 4] for i in range(n)])) / n))
# H2D running concurrently with the device-resident loop, but unsynchronised with it
cs = torch.cuda.Stream(device=dev)
stage = (torch.empty_like(dv[0][0]), torch.empty_like(dv[0][1]))


def overlapped(n):
    for i in range(n):
        with torch.cuda.stream(cs):
            stage[0].copy_(hostb[i % 4][0], non_blocking=True)
            stage[1].copy_(hostb[i % 4][1], non_blocking=True)
        ex.extract(*dv[i % 4])
    torch.cuda.current_stream().wait_stream(cs)


overlapped(10)
print('device-resident + independent H2D stream, 40 steps: %.3f ms/step' % (ev_time(lambda: overlapped(40)) / 40))
