// Implicit-GEMM convolution for sm_100a.
//
//   D[m, n] = sum_{r,s,c} X[pixel(m) + (r,s), c] * Wt[n, (r,s,c)]        m = (img, p, q), n = cout
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer : operand A by TMA *im2col* loads straight from the NHWC activation tensor
//                              (one [128 pixels x 64 channels] box per filter tap and channel chunk, zero
//                              padding and stride handled by the TMA unit), operand B (packed weights) by a
//                              tiled TMA load; both land in 128B-swizzled K-major shared memory.
//   warp 1      MMA issuer   : one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) into a double-buffered
//                              fp32 accumulator in TMEM; tcgen05.commit releases smem stages / publishes tiles.
//   warps 2..5  epilogue     : tcgen05.ld the accumulator, apply folded BatchNorm (scale, shift), residual add,
//                              per-channel PReLU/LeakyReLU slope, write bf16 channels-last (+ optional fp32
//                              side output) -- overlapped with the next tile's MMAs.
//
// Roofline: tensor pipe.  Algorithmic work = 2 * M * Cout * R*S*C flop per launch (DESIGN.md "Kernels").
#include "dl_host.cuh"
#include "dl_ptx.cuh"

namespace dl {

struct IgemmParams {
  int M, P, Q, PQ;
  int Cout;
  int cchunks, R, S;
  int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int num_m_blocks, num_n_blocks;
  int ldy, ldf;
  float f32_slope;
  const float* scale;
  const float* shift;
  const float* slope;
  const float* scale2;
  const float* shift2;
  const uint16_t* residual;
  uint16_t* y;
  float* yf;
};

constexpr int kMaxCout = 2048;   // per-channel epilogue parameters staged in shared memory

constexpr int kResidentBBytes = 72 * 1024;   // whole packed weight matrix kept in shared memory when it fits

// kResB: the weight operand (all K blocks of the single N block) is loaded once per CTA and stays resident;
// the pipeline stages then carry operand A only.  Cuts the L2 -> SM traffic of the narrow layers by a third.
template <int BLOCK_N, bool kResB>
struct IgemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int BLOCK_K = 64;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = kResB ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int STAGES = kResB ? 8 : ((BLOCK_N == 64) ? 8 : (BLOCK_N == 128 ? 6 : 4));
  static constexpr int BRES_BYTES = kResB ? kResidentBBytes : 0;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;   // two accumulator stages; 128 / 256 / 512 (power of two)
  static constexpr int BAR_BYTES = (2 * STAGES + 5) * 8 + 16;
  static constexpr int PARAM_BYTES = 3 * kMaxCout * 4;                          // scale | shift | slope
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BRES_BYTES + PARAM_BYTES + BAR_BYTES + 1024;
  static constexpr int THREADS = 320;                                           // TMA, MMA, 8 epilogue warps
};

template <int BLOCK_N, bool kResB>
__global__ void __launch_bounds__(320, 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const IgemmParams p) {
  using Cfg = IgemmCfg<BLOCK_N, kResB>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // round up inside the shared window (pointer arithmetic on the __shared__ symbol keeps LDS/STS addressing)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bres = smem + STAGES * Cfg::STAGE_BYTES;             // resident weights (kResB only)
  float* prm = reinterpret_cast<float*>(bres + Cfg::BRES_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bres + Cfg::BRES_BYTES + Cfg::PARAM_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* bfull = bars + 2 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(&tfull[0], 1);
      mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 256);
      mbar_init(&tempty[1], 256);
      mbar_init(bfull, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  if (p.y != nullptr) {
    for (int c = threadIdx.x; c < p.Cout; c += Cfg::THREADS) {
      prm[c] = p.scale[c];
      prm[kMaxCout + c] = p.shift[c];
      prm[2 * kMaxCout + c] = p.slope[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_blocks * p.num_n_blocks;
  const int num_kb = p.R * p.S * p.cchunks;

  if (warp == 0) {
    if (lane == 0) {
      if (kResB) {
        mbar_expect_tx(bfull, (uint32_t)num_kb * Cfg::B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(bres + kb * Cfg::B_BYTES, &mapB, bfull, kb * 64, 0);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_blocks;
        const int n_blk = tile - m_blk * p.num_n_blocks;
        const int m0 = m_blk * Cfg::BLOCK_M;
        const int img = m0 / p.PQ;
        const int rem = m0 - img * p.PQ;
        const int pp = rem / p.Q;
        const int qq = rem - pp * p.Q;
        const int w0 = qq * p.stride_w - p.pad_w;
        const int h0 = pp * p.stride_h - p.pad_h;
        int kb = 0;
        for (int r = 0; r < p.R; ++r) {
          for (int s = 0; s < p.S; ++s) {
            for (int cc = 0; cc < p.cchunks; ++cc, ++kb) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
              tma_load_im2col_4d(sa, &mapA, &full[stage], cc * 64, w0, h0, img, (uint16_t)(s * p.dil_w),
                                 (uint16_t)(r * p.dil_h));
              if (!kResB) tma_load_2d(sa + Cfg::A_BYTES, &mapB, &full[stage], kb * 64, n_blk * BLOCK_N);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BLOCK_N);
      if (kResB) mbar_wait(bfull, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128_kmajor(sa);
          const uint64_t bdesc = umma_desc_sw128_kmajor(kResB ? smem_u32(bres) + kb * Cfg::B_BYTES : sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 bf16 = 32 B) inside the 128-byte swizzle atom
            umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 8 warps.  Warp w reads TMEM lane
    // quarter (w & 3) -- a hardware restriction -- and every second 32-column chunk (chunk parity = (w-2) >> 2).
    const int quarter = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    const bool has_res = p.residual != nullptr;
    const bool fast = p.y != nullptr && p.yf == nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile - m_blk * p.num_n_blocks;
      const long long row = (long long)m_blk * Cfg::BLOCK_M + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int cbase = n_blk * BLOCK_N;
      // residual of this warp's first chunk: requested BEFORE waiting for the accumulator so the L2 round trip
      // overlaps the tile's MMAs
      uint4 res[4];
      {
        const int c0 = cbase + chunk0 * 32;
        if (has_res && row_ok && c0 + 32 <= p.Cout) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row * p.ldy + c0);
#pragma unroll
          for (int g = 0; g < 4; ++g) res[g] = __ldg(rp + g);
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int j = chunk0; j < BLOCK_N / 32; j += 2) {
        const int c0 = cbase + j * 32;
        if (c0 >= p.Cout) break;
        uint32_t acc_r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + j * 32, acc_r);
        const bool full_chunk = c0 + 32 <= p.Cout;
        if (fast && full_chunk) {
          // ---- straight-line path: parameters from shared memory, residual already in registers
          uint4 res_cur[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) res_cur[g] = res[g];
          if (has_res && row_ok && j + 2 < BLOCK_N / 32 && c0 + 96 <= p.Cout) {     // prefetch the next chunk's
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row * p.ldy + c0 + 64);
#pragma unroll
            for (int g = 0; g < 4; ++g) res[g] = __ldg(rp + g);
          }
          tmem_ld_wait();
          const float4* sc = reinterpret_cast<const float4*>(prm + c0);
          const float4* sh = reinterpret_cast<const float4*>(prm + kMaxCout + c0);
          const float4* sl = reinterpret_cast<const float4*>(prm + 2 * kMaxCout + c0);
          uint4 o[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 s4 = sc[2 * g + h], h4 = sh[2 * g + h];
              v[4 * h + 0] = fmaf(__uint_as_float(acc_r[8 * g + 4 * h + 0]), s4.x, h4.x);
              v[4 * h + 1] = fmaf(__uint_as_float(acc_r[8 * g + 4 * h + 1]), s4.y, h4.y);
              v[4 * h + 2] = fmaf(__uint_as_float(acc_r[8 * g + 4 * h + 2]), s4.z, h4.z);
              v[4 * h + 3] = fmaf(__uint_as_float(acc_r[8 * g + 4 * h + 3]), s4.w, h4.w);
            }
            if (has_res) {
              const uint4 rr = res_cur[g];
              v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
              v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 l4 = sl[2 * g + h];
              v[4 * h + 0] = v[4 * h + 0] > 0.f ? v[4 * h + 0] : v[4 * h + 0] * l4.x;
              v[4 * h + 1] = v[4 * h + 1] > 0.f ? v[4 * h + 1] : v[4 * h + 1] * l4.y;
              v[4 * h + 2] = v[4 * h + 2] > 0.f ? v[4 * h + 2] : v[4 * h + 2] * l4.z;
              v[4 * h + 3] = v[4 * h + 3] > 0.f ? v[4 * h + 3] : v[4 * h + 3] * l4.w;
            }
            o[g].x = pack_bf16x2(v[0], v[1]); o[g].y = pack_bf16x2(v[2], v[3]);
            o[g].z = pack_bf16x2(v[4], v[5]); o[g].w = pack_bf16x2(v[6], v[7]);
          }
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.y + row * p.ldy + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g) dst[g] = o[g];
          }
        } else {
          // ---- general path: partial chunk and / or fp32 side output (heads, attention logits)
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int c = c0 + g * 8;
              if (c >= p.Cout) continue;
              float a[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) a[i] = __uint_as_float(acc_r[g * 8 + i]);
              if (p.yf != nullptr) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  o[i] = p.scale2 != nullptr ? fmaf(a[i], __ldg(p.scale2 + c + i), __ldg(p.shift2 + c + i)) : a[i];
                  o[i] = o[i] > 0.f ? o[i] : o[i] * p.f32_slope;
                }
                float4* dst = reinterpret_cast<float4*>(p.yf + row * p.ldf + c);
                dst[0] = make_float4(o[0], o[1], o[2], o[3]);
                dst[1] = make_float4(o[4], o[5], o[6], o[7]);
              }
              if (p.y != nullptr) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaf(a[i], prm[c + i], prm[kMaxCout + c + i]);
                if (has_res) {
                  const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ldy + c));
                  v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                  v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * prm[2 * kMaxCout + c + i];
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(p.y + row * p.ldy + c) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BLOCK_N, bool kResB>
static int launch_igemm(const CUtensorMap& mapA, const CUtensorMap& mapB, const IgemmParams& p,
                        cudaStream_t stream) {
  using Cfg = IgemmCfg<BLOCK_N, kResB>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared-memory budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_conv_kernel<BLOCK_N, kResB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "igemm smem attribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  int grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (tiles < grid) grid = tiles;
  igemm_conv_kernel<BLOCK_N, kResB><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(mapA, mapB, p);
  return check_launch("igemm_conv_kernel");
}

}  // namespace dl

extern "C" int dl_conv_igemm_bf16(const void* x, const void* w_packed, const float* scale, const float* shift,
                                  const float* slope, const void* residual, void* y, float* y_f32,
                                  const float* scale2, const float* shift2, const dl_conv_desc* d, void* stream) {
  using namespace dl;
  DL_CHECK_ARG(x && w_packed && d, "conv_igemm: null x / w / desc");
  DL_CHECK_ARG(y || y_f32, "conv_igemm: no output requested");
  DL_CHECK_ARG(!y || (scale && shift && slope), "conv_igemm: bf16 output needs scale/shift/slope");
  DL_CHECK_ARG(!residual || y, "conv_igemm: residual needs the bf16 output");
  DL_CHECK_ARG((scale2 == nullptr) == (shift2 == nullptr), "conv_igemm: scale2/shift2 must come together");
  DL_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->Cout > 0, "conv_igemm: empty shape");
  DL_CHECK_ARG(d->ldx % 8 == 0 && d->ldx >= d->C, "conv_igemm: ldx must be >= C and a multiple of 8");
  DL_CHECK_ARG(d->Cout % 8 == 0 && d->Cout <= kMaxCout, "conv_igemm: Cout must be a multiple of 8, at most %d", kMaxCout);
  DL_CHECK_ARG(!y || (d->ldy % 8 == 0 && d->ldy >= d->Cout), "conv_igemm: ldy must be >= Cout, multiple of 8");
  DL_CHECK_ARG(!y_f32 || (d->ldf % 4 == 0 && d->ldf >= d->Cout), "conv_igemm: ldf must be >= Cout, multiple of 4");
  DL_CHECK_ARG(d->R >= 1 && d->S >= 1 && d->stride_h >= 1 && d->stride_w >= 1 && d->dil_h >= 1 && d->dil_w >= 1 &&
                   d->pad_h >= 0 && d->pad_w >= 0,
               "conv_igemm: bad filter geometry");
  DL_CHECK_ARG((d->S - 1) * d->dil_w <= 255 && (d->R - 1) * d->dil_h <= 255 && d->pad_w <= 127 && d->pad_h <= 127,
               "conv_igemm: filter extent exceeds the TMA im2col offset range");
  const int P = (d->H + 2 * d->pad_h - d->dil_h * (d->R - 1) - 1) / d->stride_h + 1;
  const int Q = (d->W + 2 * d->pad_w - d->dil_w * (d->S - 1) - 1) / d->stride_w + 1;
  DL_CHECK_ARG(P > 0 && Q > 0, "conv_igemm: input smaller than the filter");
  int st = require_sm100();
  if (st != DL_OK) return st;

  const long long M = (long long)d->N * P * Q;
  DL_CHECK_ARG(M < (1ll << 31) - 256, "conv_igemm: too many output pixels");
  IgemmParams p;
  p.M = (int)M; p.P = P; p.Q = Q; p.PQ = P * Q;
  p.Cout = d->Cout;
  p.cchunks = (d->C + 63) / 64;
  p.R = d->R; p.S = d->S;
  p.stride_h = d->stride_h; p.stride_w = d->stride_w;
  p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  p.dil_h = d->dil_h; p.dil_w = d->dil_w;
  p.ldy = d->ldy; p.ldf = d->ldf;
  p.f32_slope = d->f32_slope;
  p.scale = scale; p.shift = shift; p.slope = slope; p.scale2 = scale2; p.shift2 = shift2;
  p.residual = static_cast<const uint16_t*>(residual);
  p.y = static_cast<uint16_t*>(y);
  p.yf = y_f32;
  p.num_m_blocks = (int)((M + 127) / 128);

  const int block_n = d->Cout <= 64 ? 64 : (d->Cout <= 128 ? 128 : 256);
  p.num_n_blocks = (d->Cout + block_n - 1) / block_n;
  const long long Ktot = (long long)d->R * d->S * p.cchunks * 64;

  CUtensorMap mapA, mapB;
  const int img_rows = d->img_rows > 0 ? d->img_rows : d->H;
  DL_CHECK_ARG(img_rows >= d->H, "conv_igemm: img_rows < H");
  st = make_im2col_nhwc_bf16(&mapA, x, d->N, d->H, d->W, d->C, d->ldx, img_rows, d->R, d->S, d->stride_h, d->stride_w, d->pad_h,
                             d->pad_w, d->dil_h, d->dil_w, 64, 128);
  if (st != DL_OK) return st;
  st = make_tiled_2d_bf16(&mapB, w_packed, (uint64_t)d->Cout, (uint64_t)Ktot, (uint64_t)Ktot, (uint32_t)block_n, 64);
  if (st != DL_OK) return st;

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long num_kb = (long long)d->R * d->S * p.cchunks;
  const bool resident = p.num_n_blocks == 1 && num_kb * block_n * 128 <= kResidentBBytes;
  switch (block_n) {
    case 64: return resident ? launch_igemm<64, true>(mapA, mapB, p, s) : launch_igemm<64, false>(mapA, mapB, p, s);
    case 128: return resident ? launch_igemm<128, true>(mapA, mapB, p, s) : launch_igemm<128, false>(mapA, mapB, p, s);
    default: return resident ? launch_igemm<256, true>(mapA, mapB, p, s) : launch_igemm<256, false>(mapA, mapB, p, s);
  }
}
