"""CPU emulation of stem2_conv3d.cuh's data layout and index arithmetic (no GPU): pre-pass frames -> 14-row strips ->
unfolded two-plane stages -> no-swizzle K-major operand views (SBO = 128 B, LBO = plane) x the weight stack
[0, W4, W3, W2, W1, W0, 0] -> accumulator [128 lanes x 176 columns] -> per-thread BN/PReLU/3x3-s2 max-pool with the
carried conv row.  Compared with torch's conv3d + max_pool3d.  Checks the DESIGN, not the CUDA code."""
import numpy as np
import torch
import torch.nn.functional as F

WO, N, ROWB, PLANE = 44, 176, 44 * 16, 7 * 44 * 16
T, H, W = 5, 88, 88
rng = np.random.default_rng(0)
x = rng.standard_normal((T, H, W)).astype(np.float32)
w = rng.standard_normal((64, 5, 7, 7)).astype(np.float32) * 0.1
scale = rng.standard_normal(64).astype(np.float32)
shift = rng.standard_normal(64).astype(np.float32)
slope = np.abs(rng.standard_normal(64)).astype(np.float32) * 0.3

# packed weights (64, 320): K = kt*64 + kh*8 + kw (zero padding for kh = 7 / kw = 7) -- the PRODUCT's packing routine
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplip_b200 import packing
wb = torch.from_numpy(w).to(torch.bfloat16).float()              # the kernel multiplies bf16 weights
w = wb.numpy()
wp = packing.pack_stem_weight(wb[:, None]).float().numpy()
stack = np.zeros((7 * 64, 64), np.float32)
for i in range(1, 6):
    kt = 5 - i
    stack[i * 64:(i + 1) * 64] = wp[:, kt * 64:(kt + 1) * 64]

# pre-pass frames: row iy+3, col ix+3, pitch 96
xp = np.zeros((T, H + 8, 96), np.float32)
xp[:, 3:3 + H, 3:3 + W] = x

def unit(t0):
    out = np.zeros((2, 22, 22, 64), np.float32)
    carry = np.full((128, 2, 11), -np.inf, np.float32)          # [lane][half][i]
    for tile in range(H // 8):
        D = np.zeros((128, N), np.float64)
        Us = []
        for st in range(6):
            f = t0 + st - 2
            strip = xp[f, 8 * tile:8 * tile + 14] if 0 <= f < T else np.zeros((14, 96), np.float32)
            U = np.zeros(2 * PLANE // 2, np.float32)            # element (2-byte) addressed
            for c in range(14 * 48):                            # 48 work slots per strip row, 4 idle
                rr, ox = divmod(c, 48)
                if ox >= WO:
                    continue
                words = strip[rr, 2 * ox:2 * ox + 8]            # byte offset rr*192 + ox*4, 16 bytes
                a = ((rr & 1) * PLANE + (rr >> 1) * ROWB + ox * 16) // 2
                U[a:a + 8] = words
            Us.append(U)
            for j in range(3):                                  # window rows (0,1), (2,3), (4,5)
                start = j * ROWB
                Bm = np.zeros((N, 16), np.float32)
                for n in range(N):
                    for k in range(16):
                        addr = start + (n // 8) * 128 + (n % 8) * 16 + (k // 8) * PLANE + (k % 8) * 2
                        Bm[n, k] = U[addr // 2]
                A = stack[(5 - st) * 64:(5 - st) * 64 + 128, 16 * j:16 * j + 16]
                D += A.astype(np.float64) @ Bm.T.astype(np.float64)
            if st & 1:                                          # window row 6 of stages st-1, st: LBO = one stage
                Uc = np.concatenate([Us[st - 1], Us[st]])
                Bm = np.zeros((N, 16), np.float32)
                for n in range(N):
                    for k in range(16):
                        addr = 3 * ROWB + (n // 8) * 128 + (n % 8) * 16 + (k // 8) * (2 * PLANE) + (k % 8) * 2
                        Bm[n, k] = Uc[addr // 2]
                A = np.zeros((128, 16), np.float32)
                for r in range(128):
                    for c in (st - 1, st):
                        kt = c - (r >> 6)
                        if 0 <= kt <= 4:
                            A[r, 8 * (c - (st - 1)):8 * (c - (st - 1)) + 8] = wp[r & 63, kt * 64 + 48:kt * 64 + 56]
                D += A.astype(np.float64) @ Bm.T.astype(np.float64)
        # epilogue
        for lane in range(128):
            g, ch = lane // 64, lane % 64
            for half in range(2):
                col0 = 20 if half else 0
                hp = []
                for r in range(4):
                    v = D[lane, r * WO + col0:r * WO + col0 + 24]
                    z = v * scale[ch] + shift[ch]
                    z = np.where(z > 0, z, z * slope[ch])
                    if half == 0:
                        h = [max(z[0], z[1])] + [max(z[2 * i - 1], z[2 * i], z[2 * i + 1]) for i in range(1, 11)]
                    else:
                        h = [max(z[2 * i + 1], z[2 * i + 2], z[2 * i + 3]) for i in range(11)]
                    hp.append(np.array(h))
                pa = np.maximum(np.maximum(carry[lane, half], hp[0]), hp[1])
                pb = np.maximum(np.maximum(hp[1], hp[2]), hp[3])
                carry[lane, half] = hp[3]
                px0 = 11 if half else 0
                out[g, 2 * tile, px0:px0 + 11, ch] = pa
                out[g, 2 * tile + 1, px0:px0 + 11, ch] = pb
    return out

xt = torch.from_numpy(x)[None, None]
conv = F.conv3d(xt, torch.from_numpy(w)[:, None], stride=(1, 2, 2), padding=(2, 3, 3))
z = conv * torch.from_numpy(scale).view(1, 64, 1, 1, 1) + torch.from_numpy(shift).view(1, 64, 1, 1, 1)
z = torch.where(z > 0, z, z * torch.from_numpy(slope).view(1, 64, 1, 1, 1))
ref = F.max_pool3d(z, (1, 3, 3), (1, 2, 2), (0, 1, 1))[0].permute(1, 2, 3, 0).numpy()     # (T,22,22,64)
worst = 0.0
for t0 in (0, 2, 4):
    o = unit(t0)
    for g in range(min(2, T - t0)):
        worst = max(worst, float(np.abs(o[g] - ref[t0 + g]).max()))
print('max abs diff vs torch conv3d+BN+PReLU+maxpool:', worst)
assert worst < 1e-3
print('OK')
