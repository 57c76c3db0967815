"""NCCL / peer-access diagnostic for the multi-GPU job leg: which transport carries the all_gather of the
embedding table, and how fast.  torchrun --nproc-per-node N tools/nccl_diag.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplip_b200 import dist as D

rank, world, local = D.init_from_env('nccl')
dev = torch.device('cuda', local)
if rank == 0:
    os.system('nvidia-smi topo -m 1>&2')
    print('peer access 0->1:', torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None, file=sys.stderr)
for mb in (1, 14.5, 105.8, 512):
    per = int(mb * 1e6 / 4 / 1024 / world)
    full = torch.zeros((per * world, 1024), device=dev)
    loc = full[rank * per:(rank + 1) * per]
    loc.fill_(rank + 1)
    for _ in range(3):
        dist.all_gather_into_tensor(full, loc)
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_gather_into_tensor(full, loc); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ok = all(float(full[r * per]. mean()) == r + 1 for r in range(world))
    if rank == 0:
        ms = sorted(ts)[len(ts) // 2]
        print('all_gather %.1f MB total: %.3f ms  -> %.1f GB/s algbw  ok=%s' % (full.numel() * 4 / 1e6, ms, full.numel() * 4 / ms / 1e6, ok), file=sys.stderr)
dist.barrier(); dist.destroy_process_group()
