"""Fixed cost of a short persistent launch: the E-TDNN k=1 layer (512 -> 512) at growing batch sizes.  20 launches are
captured into ONE CUDA graph and replayed, so the CPU's enqueue cost (Python + ctypes + two tensor-map encodes, ~20 us
per call) is out of the measurement: what is left is kernel duration + the GPU's launch-to-launch gap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplip_b200 import _lib, ops, packing
DEV = 'cuda'
w = packing.pack_conv_weight(torch.randn(512, 512, 1, 1, device=DEV) * 0.05)
sc = torch.ones(512, device=DEV); sh = torch.zeros(512, device=DEV); sl = torch.full((512,), 0.2, device=DEV)
NL = 20
DBG = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.set_option('dbg', DBG)
print('-- dbg', DBG, '(2 = no stores, 4 = no epilogue)')
for B in (1, 32, 64, 256):
    x = torch.randn(B, 1, 288, 512, device=DEV).to(torch.bfloat16)
    y = torch.empty(B, 1, 288, 512, device=DEV, dtype=torch.bfloat16)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            ops.conv_igemm(x, w, 512, 512, 1, 1, (1, 1), (0, 0), (1, 1), sc, sh, sl, out=y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(NL):
            ops.conv_igemm(x, w, 512, 512, 1, 1, (1, 1), (0, 0), (1, 1), sc, sh, sl, out=y)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / (5 * NL)
    gf = 2.0 * B * 288 * 512 * 512 / 1e9
    print('tdnn k1 512->512  B=%-4d rows=%-6d %7.2f us per launch (graph replay)  %6.1f GFLOP  %6.0f TFLOP/s' % (B, B * 288, t, gf, gf / t / 1e3), flush=True)
