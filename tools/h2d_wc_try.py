"""Host-to-device bandwidth per rank with all ranks copying at once: ordinary pinned memory against write-combined
pinned memory (cudaHostAllocWriteCombined).  torchrun --nproc-per-node N tools/h2d_wc_try.py"""
import ctypes, os, sys, time
import torch, torch.distributed as dist
rank, local = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl')
NB = 50_380_800          # one step's inputs (64 utterances: u8 crops + int16 PCM)
cudart = ctypes.CDLL('libcudart.so.12') if os.path.exists('/usr/local/cuda/lib64/libcudart.so.12') else ctypes.CDLL('libcudart.so')
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    return t, p
dev = torch.empty(NB, dtype=torch.uint8, device='cuda')
res = {}
for name, flags in (('pinned', 0), ('write-combined', 4), ('pinned (torch)', None)):
    if flags is None:
        h = torch.empty(NB, dtype=torch.uint8).pin_memory()
    else:
        h, keep = host_alloc(NB, flags)
    h.fill_(7)
    for _ in range(3):
        dev.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dev.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    gbs = 20 * NB / (a.elapsed_time(b) / 1e3) / 1e9
    t = torch.tensor([gbs], device='cuda', dtype=torch.float64)
    if world > 1:
        lo = t.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(t)
        res[name] = (float(t), float(lo))
    else:
        res[name] = (gbs, gbs)
    if world > 1:
        dist.barrier()
if rank == 0:
    for k, (tot, lo) in res.items():
        print('%-16s aggregate %.1f GB/s over %d ranks (slowest rank %.1f GB/s)' % (k, tot, world, lo), flush=True)
