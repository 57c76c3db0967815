// K4 kernels: per-frame spatial mean + masked temporal mean in one kernel (frame_pool_kernel), and the temporal half on
// f32 frame features that already exist (temporal_mean_kernel: dl_conv_desc.avgpool took the spatial mean in the conv's
// epilogue).  Free of CUDA-runtime dependencies so that tests/frontend_cpu_emul.cpp can run this source on CPU threads.
// The includer provides: bf16_lo / bf16_hi, __ldg, uint4, float4 / make_float4, max / min.
#pragma once
#include <stdint.h>
#ifndef DL_STATIC_SHARED
#define DL_STATIC_SHARED __shared__
#endif

namespace dl {

// ------------------------------------------------------------------------------------------------
// Per-frame spatial mean, then mean over the valid frames of each utterance.  Block = (utterance, 64-channel
// slab): 8 channel-threads (8 channels = 16 B each) x 32 frame groups; deterministic smem reduction.
// kHW > 0: the map size is known at compile time and all its loads are issued before the first add (a 3x3 map is nine
// independent 16-byte loads in flight per thread instead of one); kHW = 0: any HW.  Same summation order either way.
template <int kHW>
__global__ void __launch_bounds__(256) frame_pool_kernel(const uint16_t* __restrict__ x, int T, int HW, int C,
                                                         const int32_t* __restrict__ lengths,
                                                         float* __restrict__ frame_feats,
                                                         float* __restrict__ utt_mean) {
  DL_STATIC_SHARED float sm[32][64];
  const int b = blockIdx.x;
  const int cb = blockIdx.y * 64;
  const int ct = threadIdx.x & 7, g = threadIdx.x >> 3;
  const int c = cb + ct * 8;
  int len = lengths ? lengths[b] : T;
  len = max(1, min(len, T));
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const float inv_hw = 1.f / (float)HW;
  if (c < C) {
    for (int t = g; t < T; t += 32) {
      const uint16_t* xf = x + ((size_t)(b * T + t) * HW) * C + c;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (kHW > 0) {
        uint4 vv[kHW > 0 ? kHW : 1];
#pragma unroll
        for (int px = 0; px < kHW; ++px) vv[px] = __ldg(reinterpret_cast<const uint4*>(xf + (size_t)px * C));
#pragma unroll
        for (int px = 0; px < kHW; ++px) {
          const uint4 v = vv[px];
          f[0] += bf16_lo(v.x); f[1] += bf16_hi(v.x); f[2] += bf16_lo(v.y); f[3] += bf16_hi(v.y);
          f[4] += bf16_lo(v.z); f[5] += bf16_hi(v.z); f[6] += bf16_lo(v.w); f[7] += bf16_hi(v.w);
        }
      } else {
        for (int px = 0; px < HW; ++px) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(xf + (size_t)px * C));
          f[0] += bf16_lo(v.x); f[1] += bf16_hi(v.x); f[2] += bf16_lo(v.y); f[3] += bf16_hi(v.y);
          f[4] += bf16_lo(v.z); f[5] += bf16_hi(v.z); f[6] += bf16_lo(v.w); f[7] += bf16_hi(v.w);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= inv_hw;
      if (frame_feats) {
        float4* o = reinterpret_cast<float4*>(frame_feats + ((size_t)b * T + t) * C + c);
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
      }
      if (t < len) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[g][ct * 8 + i] = acc[i];
  __syncthreads();
  if (utt_mean && threadIdx.x < 64 && cb + threadIdx.x < C) {
    float s = 0.f;
    for (int gg = 0; gg < 32; ++gg) s += sm[gg][threadIdx.x];
    utt_mean[(size_t)b * C + cb + threadIdx.x] = s / (float)len;
  }
}

// The temporal half of frame_pool_kernel on f32 frame features (the conv epilogue already took the spatial mean,
// dl_conv_desc.avgpool): same thread layout, same order of additions -> same bits as the one-kernel form.
__global__ void __launch_bounds__(256) temporal_mean_kernel(const float* __restrict__ ff, int T, int C,
                                                            const int32_t* __restrict__ lengths,
                                                            float* __restrict__ utt_mean) {
  DL_STATIC_SHARED float sm[32][64];
  const int b = blockIdx.x;
  const int cb = blockIdx.y * 64;
  const int ct = threadIdx.x & 7, g = threadIdx.x >> 3;
  const int c = cb + ct * 8;
  int len = lengths ? lengths[b] : T;
  len = max(1, min(len, T));
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (c < C) {
    for (int t = g; t < len; t += 32) {
      const float4* src = reinterpret_cast<const float4*>(ff + ((size_t)b * T + t) * C + c);
      const float4 u = __ldg(src), w = __ldg(src + 1);
      acc[0] += u.x; acc[1] += u.y; acc[2] += u.z; acc[3] += u.w;
      acc[4] += w.x; acc[5] += w.y; acc[6] += w.z; acc[7] += w.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[g][ct * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64 && cb + threadIdx.x < C) {
    float s = 0.f;
    for (int gg = 0; gg < 32; ++gg) s += sm[gg][threadIdx.x];
    utt_mean[(size_t)b * C + cb + threadIdx.x] = s / (float)len;
  }
}

}  // namespace dl
