// K7 / F1 heads: y = act(x W^T) for a handful of rows (M <= 128: one row per utterance), e.g. fc1 3000 -> 512 and
// fc2 512 -> 512 of SpeakerEmbNet.extract_embedding (models/audio_models/tdnn.py:89-101) or Linearfusion
// (models/fusion_models/model_fusion.py:19-24).  A 64 x 512 x 3000 product is 0.2 GFLOP: on the tensor-core igemm
// it occupies two CTAs for 30 us of pure latency.  Here a block owns 4 output channels, its 4 warps a quarter of K
// each and every lane RPL rows: weight chunks stream through the warp as broadcast 16-byte loads, each lane walks its
// own rows of x (one 16-byte chunk feeds 4 channels), there is no cross-lane reduction and the summation order is
// fixed (bitwise reproducible).  Same epilogue contract as the igemm kernels.
// Free of CUDA-runtime dependencies so that tests/frontend_cpu_emul.cpp can run this source on CPU threads.
// The includer provides: bf16_lo / bf16_hi (uint32 -> float), pack_bf16x2, __ldg, uint4, min.
#pragma once
#include <stdint.h>

namespace dl {

struct LinearSmallParams {
  const uint16_t* x;      // (M, ldx) bf16
  const uint16_t* w;      // (Cout, ldw) bf16, K-major (packing.py: ldw = ceil64(C))
  int M, C, ldx, ldw, Cout;
  const float* scale;     // bf16 output: v = acc * scale[c] + shift[c]; y = v > 0 ? v : v * slope[c]
  const float* shift;
  const float* slope;
  uint16_t* y;            // (M, ldy) bf16 or NULL
  int ldy;
  const float* scale2;    // f32 side output: o = acc * scale2[c] + shift2[c] (NULL: raw acc); lrelu(o, f32_slope)
  const float* shift2;
  float f32_slope;
  float* yf;              // (M, ldf) f32 or NULL
  int ldf;
};

constexpr int kLinCh = 4;       // output channels per warp (and per block)
constexpr int kLinKs = 16;      // K splits = warps per block; partial sums meet in shared memory in a fixed order

constexpr int linear_small_smem_bytes(int rpl) { return kLinKs * kLinCh * 32 * rpl * 4; }

// grid (ceil(Cout / 4), ceil(M / (32 RPL))); block 512 = 16 warps, warp ks accumulates K chunks [ks * cpk, (ks + 1) * cpk)
// of the block's 4 channels for RPL rows per lane; x chunks are loaded once per lane and reused for the 4 channels.
// The loop is a chain of L2 round trips, so its length (K / 8 / 16 chunks per warp) is what the kernel costs.
template <int RPL>
__global__ void __launch_bounds__(32 * kLinKs) linear_small_kernel(LinearSmallParams p) {
  extern __shared__ __align__(16) float lin_part[];      // [kLinKs][kLinCh][32 RPL] (dynamic: linear_small_smem_bytes)
  constexpr int kRows = 32 * RPL;
  const int ks = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kLinCh;
  const int m0 = blockIdx.y * kRows;
  const uint16_t* wr[kLinCh];
#pragma unroll
  for (int j = 0; j < kLinCh; ++j) wr[j] = p.w + (size_t)(c0 + j < p.Cout ? c0 + j : c0) * p.ldw;
  const uint16_t* xr[RPL];
  float acc[RPL][kLinCh];
#pragma unroll
  for (int r = 0; r < RPL; ++r) {
    const int m = m0 + lane + 32 * r;
    xr[r] = p.x + (size_t)(m < p.M ? m : 0) * p.ldx;      // rows beyond M compute on row 0 and are not stored
#pragma unroll
    for (int j = 0; j < kLinCh; ++j) acc[r][j] = 0.f;
  }
  const int chunks = p.C >> 3;                              // C % 8 == 0 (checked by the host)
  const int cpk = (chunks + kLinKs - 1) / kLinKs;
  const int k0 = ks * cpk, k1 = min(chunks, k0 + cpk);
#pragma unroll 2
  for (int k = k0; k < k1; ++k) {
    float x8[RPL][8];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xr[r]) + k);
      x8[r][0] = bf16_lo(xv.x); x8[r][1] = bf16_hi(xv.x); x8[r][2] = bf16_lo(xv.y); x8[r][3] = bf16_hi(xv.y);
      x8[r][4] = bf16_lo(xv.z); x8[r][5] = bf16_hi(xv.z); x8[r][6] = bf16_lo(xv.w); x8[r][7] = bf16_hi(xv.w);
    }
#pragma unroll
    for (int j = 0; j < kLinCh; ++j) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wr[j]) + k);
      const float w8[8] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y),
                           bf16_lo(wv.z), bf16_hi(wv.z), bf16_lo(wv.w), bf16_hi(wv.w)};
#pragma unroll
      for (int r = 0; r < RPL; ++r) {
        float a = acc[r][j];
#pragma unroll
        for (int i = 0; i < 8; ++i) a = fmaf(x8[r][i], w8[i], a);
        acc[r][j] = a;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RPL; ++r)
#pragma unroll
    for (int j = 0; j < kLinCh; ++j) lin_part[(ks * kLinCh + j) * kRows + lane + 32 * r] = acc[r][j];
  __syncthreads();
  // ---- epilogue: thread <-> (channel, row) pairs of the block; K partials summed in a fixed order
  for (int i = threadIdx.x; i < kLinCh * 32 * RPL; i += 32 * kLinKs) {
    const int j = i / kRows, ml = i - j * kRows;
    const int c = c0 + j, m = m0 + ml;
    if (m >= p.M || c >= p.Cout) continue;
    const float* pj = lin_part + j * kRows + ml;
    float a = pj[0];
#pragma unroll
    for (int q = 1; q < kLinKs; ++q) a += pj[q * kLinCh * kRows];
    if (p.yf != nullptr) {
      float o = p.scale2 != nullptr ? fmaf(a, __ldg(p.scale2 + c), __ldg(p.shift2 + c)) : a;
      o = o > 0.f ? o : o * p.f32_slope;
      p.yf[(size_t)m * p.ldf + c] = o;
    }
    if (p.y != nullptr) {
      float v = fmaf(a, __ldg(p.scale + c), __ldg(p.shift + c));
      v = v > 0.f ? v : v * __ldg(p.slope + c);
      p.y[(size_t)m * p.ldy + c] = (uint16_t)(pack_bf16x2(v, 0.f) & 0xffffu);
    }
  }
}

}  // namespace dl
