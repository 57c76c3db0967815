// HBM-bound reduction / gather kernels of the hot path: statistics pooling (K6), per-frame average pool +
// masked temporal mean (K4), z-norm + concat + L2 fusion (K8), trial cosine scoring (K9) and the layout
// shims around them.  All are warp-shuffle reductions with 16-byte coalesced loads; shared memory is only
// used to combine partial results across warps.
#include "dl_host.cuh"
#include "dl_ptx.cuh"
#include "plda_score.cuh"
#include "pool_kernels.cuh"

namespace dl {

// ------------------------------------------------------------------------------------------------
// Statistics pooling over time, channels-last bf16 input.  Block = 8 warps; a warp reads 256 channels
// (32 lanes x 8 bf16 = 16 B per lane) of one time step per iteration; the 8 warps stride over time and
// are merged with Chan's parallel-variance update.  Optional attention weights (softmax over time).
template <bool kAttn, int kMlp, int kSlab>
__global__ void __launch_bounds__(256) stat_pool_kernel(const uint16_t* __restrict__ x, const float* __restrict__ logits,
                                                        int T, int C, int ldx, const int32_t* __restrict__ lengths,
                                                        float* __restrict__ out_f32, uint16_t* __restrict__ out_bf16,
                                                        int ld_out) {
  // kSlab channels per block (256: a warp per time step, 8 "virtual warps" stride over time; 128: a half warp per
  // time step, 16 virtual warps -- twice the blocks for the same bytes, better balance over 148 SMs)
  constexpr int kVW = 8 * (256 / kSlab), kLanes = kSlab / 8;
  extern __shared__ float sm[];   // kAttn: alpha[T] ; then kVW virtual warps x kSlab ch x 2 partials
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vw = threadIdx.x / kLanes, vl = threadIdx.x % kLanes;
  const int c0 = blockIdx.x * kSlab + vl * 8;
  int len = lengths ? lengths[b] : T;
  len = max(1, min(len, T));
  const uint16_t* xb = x + (size_t)b * T * ldx;

  float* alpha = sm;
  float* part = sm + (kAttn ? T : 0);
  if (kAttn) {
    // softmax over the valid time steps (models/audio_models/pooling.py:100)
    const float* e = logits + (size_t)b * T;
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < len; t += 256) mx = fmaxf(mx, e[t]);
    mx = warp_max(mx);
    __shared__ float red[8];
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int t = threadIdx.x; t < len; t += 256) {
      float w = __expf(e[t] - mx);
      alpha[t] = w;
      s += w;
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    const float inv = 1.f / s;
    for (int t = threadIdx.x; t < len; t += 256) alpha[t] *= inv;
    __syncthreads();
  }

  float a0[8], a1[8];   // kAttn: sum(alpha x), sum(alpha x^2); else Welford mean, M2
#pragma unroll
  for (int i = 0; i < 8; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
  int n = 0;
  if (c0 < C) {
    // kMlp time steps (independent 16-byte loads) in flight per lane: the kernel is pure streaming (same accumulation
    // order for every kMlp)
    for (int tb = vw; tb < len; tb += kVW * kMlp) {
      uint4 v4[kMlp];
#pragma unroll
      for (int u = 0; u < kMlp; ++u) {
        const int t = tb + kVW * u;
        if (t < len) v4[u] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)t * ldx + c0));
      }
#pragma unroll
      for (int u = 0; u < kMlp; ++u) {
        const int t = tb + kVW * u;
        if (t >= len) break;
        const uint4 v = v4[u];
        float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                      bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
        if (kAttn) {
          const float w = alpha[t];
#pragma unroll
          for (int i = 0; i < 8; ++i) { a0[i] = fmaf(w, f[i], a0[i]); a1[i] = fmaf(w * f[i], f[i], a1[i]); }
        } else {
          ++n;
          const float rn = 1.f / (float)n;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float d = f[i] - a0[i];
            a0[i] += d * rn;
            a1[i] = fmaf(d, f[i] - a0[i], a1[i]);
          }
        }
      }
    }
  }
  // combine the virtual warps
  float* p0 = part + (vw * kSlab + vl * 8) * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) { p0[2 * i] = a0[i]; p0[2 * i + 1] = a1[i]; }
  __shared__ int cnt[kVW];
  if (vl == 0) cnt[vw] = n;
  __syncthreads();
  const int cl = threadIdx.x;          // one thread per channel of this block's slab
  const int c = blockIdx.x * kSlab + cl;
  if (cl < kSlab && c < C) {
    float mean, sd;
    if (kAttn) {
      float s1 = 0.f, s2 = 0.f;
      for (int w = 0; w < kVW; ++w) { s1 += part[(w * kSlab + cl) * 2]; s2 += part[(w * kSlab + cl) * 2 + 1]; }
      mean = s1;
      sd = sqrtf(s2 - s1 * s1);        // no clamp, like the reference (pooling.py:105)
    } else {
      float m = 0.f, m2 = 0.f;
      int nn = 0;
      for (int w = 0; w < kVW; ++w) {
        const int nw = cnt[w];
        if (nw == 0) continue;
        const float mw = part[(w * kSlab + cl) * 2], m2w = part[(w * kSlab + cl) * 2 + 1];
        const int nt = nn + nw;
        const float d = mw - m;
        m += d * ((float)nw / (float)nt);
        m2 += m2w + d * d * ((float)nn * (float)nw / (float)nt);
        nn = nt;
      }
      mean = m;
      sd = sqrtf(m2 / (float)(nn - 1));   // unbiased (torch.std default); len==1 -> nan like torch
    }
    if (out_f32) {
      out_f32[(size_t)b * 2 * C + c] = mean;
      out_f32[(size_t)b * 2 * C + C + c] = sd;
    }
    if (out_bf16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16) + (size_t)b * ld_out;
      o[c] = __float2bfloat16_rn(mean);
      o[C + c] = __float2bfloat16_rn(sd);
    }
  }
}

// e[row] = sum_h v[h] * relu(hf[row, h]) + k ; one warp per row.
__global__ void attn_logits_kernel(const float* __restrict__ hf, int rows, int Hd, int ldh,
                                   const float* __restrict__ v, float k, float* __restrict__ e) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int h = lane; h < Hd; h += 32) s = fmaf(__ldg(v + h), fmaxf(hf[(size_t)row * ldh + h], 0.f), s);
  s = warp_sum(s);
  if (lane == 0) e[row] = s + k;
}

// K4: frame_pool_kernel (spatial mean + masked temporal mean) and temporal_mean_kernel (the temporal half alone) live in
// pool_kernels.cuh, which tests/frontend_cpu_emul.cpp also compiles for the CPU.

// ------------------------------------------------------------------------------------------------
// Row z-norm of two modalities + concat (+ L2).  One warp per utterance.
__device__ __forceinline__ void row_stats(const float* r, int D, int lane, bool biased, float& mean, float& sd) {
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += r[i];
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) { const float d = r[i] - mean; q = fmaf(d, d, q); }
  q = warp_sum(q);
  sd = sqrtf(q / (float)(biased ? D : D - 1));
}

__global__ void znorm_concat_kernel(const float* __restrict__ a, int Da, const float* __restrict__ v, int Dv, int B,
                                    int biased, int video_first, int l2norm, float* __restrict__ out,
                                    uint16_t* __restrict__ out_bf16) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const float* ra = a + (size_t)row * Da;
  const float* rv = v + (size_t)row * Dv;
  float ma, sa, mv, sv;
  row_stats(ra, Da, lane, biased != 0, ma, sa);
  row_stats(rv, Dv, lane, biased != 0, mv, sv);
  const int D = Da + Dv;
  const int off_a = video_first ? Dv : 0, off_v = video_first ? 0 : Da;
  float nrm = 1.f;
  if (l2norm) {
    float q = 0.f;
    for (int i = lane; i < Da; i += 32) { const float z = (ra[i] - ma) / sa; q = fmaf(z, z, q); }
    for (int i = lane; i < Dv; i += 32) { const float z = (rv[i] - mv) / sv; q = fmaf(z, z, q); }
    q = sqrtf(warp_sum(q));
    nrm = q == 0.f ? 1.f : q;
  }
  for (int i = lane; i < Da; i += 32) {
    const float z = (ra[i] - ma) / sa / nrm;
    if (out) out[(size_t)row * D + off_a + i] = z;
    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out_bf16)[(size_t)row * D + off_a + i] = __float2bfloat16_rn(z);
  }
  for (int i = lane; i < Dv; i += 32) {
    const float z = (rv[i] - mv) / sv / nrm;
    if (out) out[(size_t)row * D + off_v + i] = z;
    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out_bf16)[(size_t)row * D + off_v + i] = __float2bfloat16_rn(z);
  }
}

__global__ void lowfer_kernel(const float* __restrict__ e1, const float* __restrict__ e2, int B, int D,
                              float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * D) return;
  const size_t r = i / D, c = i % D;
  const float x = e1[i];
  const float s = 1.f / (1.f + expf(-e2[i]));
  out[r * 3 * D + c] = x;
  out[r * 3 * D + D + c] = s;
  out[r * 3 * D + 2 * D + c] = s * x;
}

__global__ void l2_normalize_kernel(const float* __restrict__ x, int B, int D, float* __restrict__ out,
                                    uint16_t* __restrict__ out_bf16) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  const float* r = x + (size_t)row * D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) q = fmaf(r[i], r[i], q);
  q = sqrtf(warp_sum(q));
  const float n = q == 0.f ? 1.f : q;
  for (int i = lane; i < D; i += 32) {
    const float z = r[i] / n;
    if (out) out[(size_t)row * D + i] = z;
    if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out_bf16)[(size_t)row * D + i] = __float2bfloat16_rn(z);
  }
}

__global__ void affine_act_kernel(const float* __restrict__ x, int rows, int C, const float* __restrict__ scale,
                                  const float* __restrict__ shift, float slope, uint16_t* __restrict__ y, int ldc,
                                  float* __restrict__ yf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * ldc) return;
  const size_t r = i / ldc;
  const int c = (int)(i % ldc);
  float v = 0.f;
  if (c < C) {
    v = x[r * C + c];
    if (scale) v = fmaf(v, scale[c], shift[c]);
    v = v > 0.f ? v : v * slope;
    if (yf) yf[r * C + c] = v;
  }
  if (y) reinterpret_cast<__nv_bfloat16*>(y)[i] = __float2bfloat16_rn(v);
}

// (B, C, T) f32 -> (B, T, ldc) bf16, zero padded channels; 32x32 smem transpose tiles.
__global__ void nct_to_ntc_kernel(const float* __restrict__ x, int C, int T, uint16_t* __restrict__ y, int ldc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < T) ? x[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < ldc)
      reinterpret_cast<__nv_bfloat16*>(y)[((size_t)b * T + t) * ldc + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Trial scoring, gather formulation: one warp per trial, two embedding rows streamed with 16-byte loads.
// Algorithmic bytes: every unique embedding once + 12 B per trial (DESIGN.md "K9").
__device__ __forceinline__ void pair_dot(const float* __restrict__ a, const float* __restrict__ b, int D, int lane,
                                         float& dot, float& na, float& nb) {
  float d = 0.f, qa = 0.f, qb = 0.f;
  if ((D & 3) == 0) {
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll 8
    for (int i = lane; i < D / 4; i += 32) {
      const float4 u = __ldg(a4 + i), w = __ldg(b4 + i);
      d = fmaf(u.x, w.x, d); d = fmaf(u.y, w.y, d); d = fmaf(u.z, w.z, d); d = fmaf(u.w, w.w, d);
      qa = fmaf(u.x, u.x, qa); qa = fmaf(u.y, u.y, qa); qa = fmaf(u.z, u.z, qa); qa = fmaf(u.w, u.w, qa);
      qb = fmaf(w.x, w.x, qb); qb = fmaf(w.y, w.y, qb); qb = fmaf(w.z, w.z, qb); qb = fmaf(w.w, w.w, qb);
    }
  } else {
    for (int i = lane; i < D; i += 32) {
      const float u = a[i], w = b[i];
      d = fmaf(u, w, d); qa = fmaf(u, u, qa); qb = fmaf(w, w, qb);
    }
  }
  dot = warp_sum(d);
  na = sqrtf(warp_sum(qa));
  nb = sqrtf(warp_sum(qb));
}

__global__ void __launch_bounds__(256) cosine_trials_kernel(const float* __restrict__ emb, int n_utt, int D,
                                                            const int32_t* __restrict__ enrol,
                                                            const int32_t* __restrict__ test, int n_trials,
                                                            float* __restrict__ scores) {
  const int trial = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (trial >= n_trials) return;
  const int lane = threadIdx.x & 31;
  const int i = enrol[trial], j = test[trial];
  if (i < 0 || i >= n_utt || j < 0 || j >= n_utt) {
    if (lane == 0) scores[trial] = __int_as_float(0x7fc00000);
    return;
  }
  float dot, na, nb;
  pair_dot(emb + (size_t)i * D, emb + (size_t)j * D, D, lane, dot, na, nb);
  // sklearn normalises rows first; a zero row is divided by 1 (models/fusion_models/utils.py:278)
  if (na == 0.f) na = 1.f;
  if (nb == 0.f) nb = 1.f;
  if (lane == 0) scores[trial] = dot / (na * nb);
}

__global__ void __launch_bounds__(256) score_fusion_kernel(const float* __restrict__ ea, int Da,
                                                           const float* __restrict__ ev, int Dv, int n_utt,
                                                           const int32_t* __restrict__ enrol,
                                                           const int32_t* __restrict__ test, int n_trials,
                                                           float* __restrict__ scores) {
  const int trial = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (trial >= n_trials) return;
  const int lane = threadIdx.x & 31;
  const int i = enrol[trial], j = test[trial];
  if (i < 0 || i >= n_utt || j < 0 || j >= n_utt) {
    if (lane == 0) scores[trial] = __int_as_float(0x7fc00000);
    return;
  }
  float dot, na, nb;
  pair_dot(ea + (size_t)i * Da, ea + (size_t)j * Da, Da, lane, dot, na, nb);
  if (na == 0.f) na = 1.f;
  if (nb == 0.f) nb = 1.f;
  const float sa = dot / (na * nb);
  pair_dot(ev + (size_t)i * Dv, ev + (size_t)j * Dv, Dv, lane, dot, na, nb);
  // F.cosine_similarity(..., eps=1e-8): each norm clamped from below (utils.py:372)
  const float sv = dot / (fmaxf(na, 1e-8f) * fmaxf(nb, 1e-8f));
  if (lane == 0) scores[trial] = 0.5f * sa + 0.5f * sv;
}

__global__ void gather_scores_kernel(const float* __restrict__ S, int ld, const int32_t* __restrict__ rows,
                                     const int32_t* __restrict__ cols, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = S[(size_t)rows[i] * ld + cols[i]];
}

}  // namespace dl

// ================================================================================================ C ABI
using namespace dl;

extern "C" int dl_stat_pool(const void* x, int B, int T, int C, int ldx, const int32_t* lengths, float* out_f32,
                            void* out_bf16, int ld_out, void* stream) {
  DL_CHECK_ARG(x && (out_f32 || out_bf16), "stat_pool: null pointer");
  DL_CHECK_ARG(B > 0 && T > 0 && C > 0 && ldx % 8 == 0 && ldx >= C, "stat_pool: bad shape B=%d T=%d C=%d ldx=%d", B,
               T, C, ldx);
  DL_CHECK_ARG(!out_bf16 || ld_out >= 2 * C, "stat_pool: ld_out < 2C");
  const size_t smem = 8 * 256 * 2 * sizeof(float);
  if (opt_statpool_slab() == 128) {
    stat_pool_kernel<false, 4, 128><<<dim3((C + 127) / 128, B), 256, smem, (cudaStream_t)stream>>>(
        (const uint16_t*)x, nullptr, T, C, ldx, lengths, out_f32, (uint16_t*)out_bf16, ld_out);
    return check_launch("stat_pool_kernel");
  }
  dim3 grid((C + 255) / 256, B);
  if (opt_statpool_mlp() == 4)
    stat_pool_kernel<false, 4, 256><<<grid, 256, smem, (cudaStream_t)stream>>>(
        (const uint16_t*)x, nullptr, T, C, ldx, lengths, out_f32, (uint16_t*)out_bf16, ld_out);
  else
    stat_pool_kernel<false, 8, 256><<<grid, 256, smem, (cudaStream_t)stream>>>(
        (const uint16_t*)x, nullptr, T, C, ldx, lengths, out_f32, (uint16_t*)out_bf16, ld_out);
  return check_launch("stat_pool_kernel");
}

extern "C" int dl_attn_stat_pool(const void* x, const float* logits, int B, int T, int C, int ldx,
                                 const int32_t* lengths, float* out_f32, void* out_bf16, int ld_out, void* stream) {
  DL_CHECK_ARG(x && logits && (out_f32 || out_bf16), "attn_stat_pool: null pointer");
  DL_CHECK_ARG(B > 0 && T > 0 && T <= 8192 && C > 0 && ldx % 8 == 0 && ldx >= C, "attn_stat_pool: bad shape");
  DL_CHECK_ARG(!out_bf16 || ld_out >= 2 * C, "attn_stat_pool: ld_out < 2C");
  dim3 grid((C + 255) / 256, B);
  const size_t smem = (8 * 256 * 2 + T) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(stat_pool_kernel<true, 4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "attn_stat_pool smem: %s", cudaGetErrorString(e));
  }
  stat_pool_kernel<true, 4, 256><<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const uint16_t*)x, logits, T, C, ldx, lengths, out_f32, (uint16_t*)out_bf16, ld_out);
  return check_launch("attn_stat_pool_kernel");
}

extern "C" int dl_attn_logits(const float* h, int rows, int Hd, int ldh, const float* v, float k, float* e,
                              void* stream) {
  DL_CHECK_ARG(h && v && e && rows > 0 && Hd > 0 && ldh >= Hd, "attn_logits: bad argument");
  attn_logits_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(h, rows, Hd, ldh, v, k, e);
  return check_launch("attn_logits_kernel");
}

extern "C" int dl_frame_pool_temporal_mean(const void* x, int B, int T, int HW, int C, const int32_t* lengths,
                                           float* frame_feats, float* utt_mean, void* stream) {
  DL_CHECK_ARG(x && (frame_feats || utt_mean), "frame_pool: null pointer");
  DL_CHECK_ARG(B > 0 && T > 0 && HW > 0 && C > 0 && C % 8 == 0 && C <= 2048, "frame_pool: bad shape");
  dim3 grid(B, (C + 63) / 64);
  if (HW == 9)         // the 3x3 maps ResNet-18 leaves of an 88x88 crop
    frame_pool_kernel<9><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x, T, HW, C, lengths, frame_feats,
                                                                 utt_mean);
  else
    frame_pool_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x, T, HW, C, lengths, frame_feats,
                                                                 utt_mean);
  return check_launch("frame_pool_kernel");
}

extern "C" int dl_temporal_mean_f32(const float* frame_feats, int B, int T, int C, const int32_t* lengths,
                                    float* utt_mean, void* stream) {
  DL_CHECK_ARG(frame_feats && utt_mean, "temporal_mean: null pointer");
  DL_CHECK_ARG(B > 0 && T > 0 && C > 0 && C % 8 == 0 && C <= 2048, "temporal_mean: bad shape");
  DL_CHECK_ARG((reinterpret_cast<uintptr_t>(frame_feats) & 15) == 0, "temporal_mean: frame_feats must be 16-byte aligned");
  temporal_mean_kernel<<<dim3(B, (C + 63) / 64), 256, 0, (cudaStream_t)stream>>>(frame_feats, T, C, lengths, utt_mean);
  return check_launch("temporal_mean_kernel");
}

extern "C" int dl_znorm_concat(const float* a, int Da, const float* v, int Dv, int B, int biased, int video_first,
                               int l2norm, float* out, void* out_bf16, void* stream) {
  DL_CHECK_ARG(a && v && (out || out_bf16) && B > 0 && Da > 1 && Dv > 1, "znorm_concat: bad argument");
  znorm_concat_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(a, Da, v, Dv, B, biased, video_first, l2norm, out,
                                                                    (uint16_t*)out_bf16);
  return check_launch("znorm_concat_kernel");
}

extern "C" int dl_lowfer(const float* e1, const float* e2, int B, int D, float* out, void* stream) {
  DL_CHECK_ARG(e1 && e2 && out && B > 0 && D > 0, "lowfer: bad argument");
  const size_t n = (size_t)B * D;
  lowfer_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e1, e2, B, D, out);
  return check_launch("lowfer_kernel");
}

extern "C" int dl_l2_normalize(const float* x, int B, int D, float* out, void* out_bf16, void* stream) {
  DL_CHECK_ARG(x && (out || out_bf16) && B > 0 && D > 0, "l2_normalize: bad argument");
  l2_normalize_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, B, D, out, (uint16_t*)out_bf16);
  return check_launch("l2_normalize_kernel");
}

extern "C" int dl_affine_act(const float* x, int rows, int C, const float* scale, const float* shift, float slope,
                             void* y_bf16, int ldc, float* y_f32, void* stream) {
  DL_CHECK_ARG(x && (y_bf16 || y_f32) && rows > 0 && C > 0, "affine_act: bad argument");
  if (!y_bf16) ldc = C;
  DL_CHECK_ARG(ldc >= C, "affine_act: ldc < C");
  DL_CHECK_ARG((scale == nullptr) == (shift == nullptr), "affine_act: scale/shift must come together");
  const size_t n = (size_t)rows * ldc;
  affine_act_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, rows, C, scale, shift, slope,
                                                                                  (uint16_t*)y_bf16, ldc, y_f32);
  return check_launch("affine_act_kernel");
}

extern "C" int dl_nct_to_ntc_bf16(const float* x, int B, int C, int T, void* y, int ldc, void* stream) {
  DL_CHECK_ARG(x && y && B > 0 && C > 0 && T > 0 && ldc >= C, "nct_to_ntc: bad argument");
  dim3 grid((T + 31) / 32, (ldc + 31) / 32, B), block(32, 8);
  nct_to_ntc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, C, T, (uint16_t*)y, ldc);
  return check_launch("nct_to_ntc_kernel");
}

extern "C" int dl_cosine_score_trials(const float* emb, int n_utt, int D, const int32_t* enrol, const int32_t* test,
                                      int n_trials, float* scores, void* stream) {
  if (n_trials == 0) return DL_OK;
  DL_CHECK_ARG(emb && enrol && test && scores && n_utt > 0 && D > 0, "cosine_score: bad argument");
  DL_CHECK_ARG(n_trials > 0, "cosine_score: negative trial count");
  cosine_trials_kernel<<<(n_trials + 7) / 8, 256, 0, (cudaStream_t)stream>>>(emb, n_utt, D, enrol, test, n_trials,
                                                                            scores);
  return check_launch("cosine_trials_kernel");
}

extern "C" int dl_score_fusion_trials(const float* emb_a, int Da, const float* emb_v, int Dv, int n_utt,
                                      const int32_t* enrol, const int32_t* test, int n_trials, float* scores,
                                      void* stream) {
  if (n_trials == 0) return DL_OK;
  DL_CHECK_ARG(emb_a && emb_v && enrol && test && scores && n_utt > 0 && Da > 0 && Dv > 0, "score_fusion: bad argument");
  DL_CHECK_ARG(n_trials > 0, "score_fusion: negative trial count");
  score_fusion_kernel<<<(n_trials + 7) / 8, 256, 0, (cudaStream_t)stream>>>(emb_a, Da, emb_v, Dv, n_utt, enrol, test,
                                                                           n_trials, scores);
  return check_launch("score_fusion_kernel");
}

extern "C" int dl_gather_scores(const float* S, int ld, const int32_t* rows, const int32_t* cols, int n_trials,
                                float* scores, void* stream) {
  if (n_trials == 0) return DL_OK;
  DL_CHECK_ARG(S && rows && cols && scores && ld > 0, "gather_scores: bad argument");
  gather_scores_kernel<<<(n_trials + 255) / 256, 256, 0, (cudaStream_t)stream>>>(S, ld, rows, cols, n_trials, scores);
  return check_launch("gather_scores_kernel");
}

extern "C" int dl_plda_transform(const float* emb, int n_utt, int D, const float* M, const float* bias, int R, float* u,
                                 void* stream) {
  if (n_utt == 0) return DL_OK;
  DL_CHECK_ARG(emb && M && bias && u && n_utt > 0 && D > 0, "plda_transform: bad argument");
  DL_CHECK_ARG(R >= 1 && R <= kPldaMaxR, "plda_transform: 1 <= R <= %d relevant dimensions (got %d)", kPldaMaxR, R);
  plda_transform_kernel<<<(n_utt + 7) / 8, 256, 0, (cudaStream_t)stream>>>(emb, n_utt, D, M, bias, R, u);
  return check_launch("plda_transform_kernel");
}

extern "C" int dl_plda_llr_trials(const float* u, int n_utt, int R, const float* k1, const float* k2, float c0,
                                  const int32_t* enrol, const int32_t* test, int n_trials, float* scores, void* stream) {
  if (n_trials == 0) return DL_OK;
  DL_CHECK_ARG(u && k1 && k2 && enrol && test && scores && n_utt > 0, "plda_llr: bad argument");
  DL_CHECK_ARG(R >= 1 && R <= kPldaMaxR && n_trials > 0, "plda_llr: 1 <= R <= %d, n_trials >= 0", kPldaMaxR);
  plda_llr_trials_kernel<<<(n_trials + 255) / 256, 256, 0, (cudaStream_t)stream>>>(u, n_utt, R, k1, k2, c0, enrol, test,
                                                                                   n_trials, scores);
  return check_launch("plda_llr_trials_kernel");
}
