"""Host-to-device bandwidth per rank with all ranks copying at once, with the GPU idle and with the extraction step
running on another stream (what bench.py's e2e leg does).  torchrun --nproc-per-node N tools/h2d_under_load.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from deeplip_b200.pipeline import AVExtractor, build_models
rank, local = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
raw, wav = bench.synth_batch(64, seed=1)
hr, hw = torch.from_numpy(raw).pin_memory(), torch.from_numpy(wav).pin_memory()
dr, dw = hr.cuda(), hw.cuda()
sr, sw = torch.empty_like(dr), torch.empty_like(dw)
NB = hr.numel() * hr.element_size() + hw.numel() * hw.element_size()
copy = torch.cuda.Stream()
for _ in range(3):
    ex.extract(dw, dr)
torch.cuda.synchronize()
def measure(load, n=20):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    with torch.cuda.stream(copy):
        a.record(copy)
        for _ in range(n):
            sw.copy_(hw, non_blocking=True); sr.copy_(hr, non_blocking=True)
        b.record(copy)
    if load:
        for _ in range(n):
            ex.extract(dw, dr)
    c1.record()
    torch.cuda.synchronize()
    gbs = n * NB / (a.elapsed_time(b) / 1e3) / 1e9
    t = torch.tensor([gbs, c0.elapsed_time(c1) / n], device='cuda', dtype=torch.float64)
    if world > 1:
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        hi = t.clone(); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        return float(t[0]), float(lo[0]), float(hi[1])
    return gbs, gbs, float(t[1])
for load in (False, True, False, True):
    tot, lo, step = measure(load)
    if rank == 0:
        print('%-26s H2D aggregate %.1f GB/s over %d ranks (slowest rank %.1f GB/s)%s' % (
            'GPU running the step:' if load else 'GPU idle:', tot, world, lo, ('; step %.3f ms (slowest rank)' % step) if load else ''), flush=True)
if world > 1:
    dist.destroy_process_group()
