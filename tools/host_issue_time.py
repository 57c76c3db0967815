"""How long the HOST needs to issue one extraction step (35 launches through ctypes + the torch glue) against how long
the GPU needs to run it: eager, and as a CUDA-graph replay.   python tools/host_issue_time.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from deeplip_b200.pipeline import AVExtractor, GraphedExtractor, build_models
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))          # under torchrun: N independent ranks contending for the host
RANK = int(os.environ.get('RANK', 0))
audio, video = build_models('cuda', seed=1)
ex = AVExtractor(audio, video)
batches = []
for r in range(4):
    raw, wav = bench.synth_batch(64, seed=r + 1)
    batches.append((torch.from_numpy(wav).cuda(), torch.from_numpy(raw).cuda()))
for i in range(5):
    ex.extract(*batches[i % 4])
torch.cuda.synchronize()
if RANK == 0: print('OMP_NUM_THREADS', os.environ.get('OMP_NUM_THREADS'), 'torch threads', torch.get_num_threads(), 'cpus', len(os.sched_getaffinity(0)))
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        ex.extract(*batches[i % 4])
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if RANK == 0: print('eager: host issue %.3f ms/step, GPU %.3f ms/step, wall %.3f ms/step' % ((t1 - t0) / n * 1e3, a.elapsed_time(b) / n, (t2 - t0) / n * 1e3), flush=True)
g = GraphedExtractor(ex, *batches[0])
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for i in range(n):
        g.extract(*batches[i % 4])
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if RANK == 0: print('graph: host issue %.3f ms/step, GPU %.3f ms/step (incl. the two input copies)' % ((t1 - t0) / n * 1e3, a.elapsed_time(b) / n), flush=True)
