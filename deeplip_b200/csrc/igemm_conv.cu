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
#include "igemm_common.cuh"
#include "linear_small.cuh"

namespace dl {



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
    if (elect_one_sync()) {
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
              if (p.lin)                   // guarded-linear A: the tile's rows shifted by the tap, plain 2D box
                tma_load_2d(sa, &mapA, &full[stage], cc * 64, m0 + (r * p.dil_h - p.pad_h) * p.lin_w + s * p.dil_w - p.pad_w);
              else
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
    if (elect_one_sync()) {           // one thread, known to ptxas as such: UTCHMMA / UTCBAR issue without per-lane loops
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
      long long row = (long long)m_blk * Cfg::BLOCK_M + quarter * 32 + lane;
      const bool row_ok = igemm_map_row(p, row);
      const int cbase = n_blk * BLOCK_N;
      // residual of this warp's first chunk: requested BEFORE waiting for the accumulator so the L2 round trip
      // overlaps the tile's MMAs
      uint4 res[4];
      igemm_prefetch_residual(p, row, row_ok, cbase, chunk0, has_res, res);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      igemm_epilogue_tile<BLOCK_N>(p, prm, kMaxCout, tmem_base + acc * BLOCK_N, row, row_ok, cbase, quarter, chunk0, has_res, fast,
                                   res);
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
  static PerDevice<bool> configured_dev;
  bool* configured = configured_dev.slot();
  if (!configured) return fail(DL_ERR_CUDA, "igemm: no current device");
  if (!*configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_conv_kernel<BLOCK_N, kResB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "igemm smem attribute: %s", cudaGetErrorString(e));
    *configured = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  int grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (tiles < grid) grid = tiles;
  igemm_conv_kernel<BLOCK_N, kResB><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(mapA, mapB, p);
  return check_launch("igemm_conv_kernel");
}

int launch_igemm_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const IgemmParams& p, int block_n,
                      cudaStream_t stream, const CUtensorMap* mapY, const CUtensorMap* mapY2);   // igemm2_conv.cu
int igemm_pair_staged(const IgemmParams& p, int block_n, bool want_half);   // 0 none, 1 = 128-wide, 2 = half-skip entry tile, 3 = streaming 256-wide
int igemm_pair_taps(const IgemmParams& p, int block_n);
bool igemm_pair_resident(const IgemmParams& p, int block_n);   // the pair kernel would keep this CTA's weight half in smem

}  // namespace dl

namespace dl {
// Unfused form of dl_conv_desc.avgpool: the pooling kernel over y (N images of `pool` rows).  frame_pool_kernel works
// on (utterance, frame) pairs with 32 frame groups per block: N images are presented as N / 32 "utterances" of 32
// frames (+ a tail of single-frame ones), which keeps all 256 threads of a block busy; same arithmetic per image.
static int avgpool_unfused(const void* y, int N, int pool, int Cout, float* out, void* stream) {
  const int nb = N / 32, rem = N - nb * 32;
  int st = DL_OK;
  if (nb > 0) st = dl_frame_pool_temporal_mean(y, nb, 32, pool, Cout, nullptr, out, nullptr, stream);
  if (st == DL_OK && rem > 0)
    st = dl_frame_pool_temporal_mean(static_cast<const uint16_t*>(y) + (size_t)nb * 32 * pool * Cout, rem, 1, pool, Cout, nullptr,
                                     out + (size_t)nb * 32 * Cout, nullptr, stream);
  return st;
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
  DL_CHECK_ARG(d->Cout % 8 == 0, "conv_igemm: Cout must be a multiple of 8 (pad the packed weights)");
  DL_CHECK_ARG(!y || d->Cout <= kMaxCout, "conv_igemm: the bf16 epilogue stages at most %d channels", kMaxCout);
  const int split = d->split_channel;
  if (split > 0) {
    DL_CHECK_ARG(y && d->y_split && !residual && !y_f32 && !d->lin && d->out_img_rows == 0,
                 "conv_igemm: split output needs y and y_split, no residual / f32 output / guarded layouts");
    DL_CHECK_ARG(split % 256 == 0 && split < d->Cout && d->Cout > 128, "conv_igemm: split_channel must be a multiple of 256 below Cout");
    DL_CHECK_ARG(d->ldy % 8 == 0 && d->ldy >= split && d->ldy >= d->Cout - split, "conv_igemm: ldy too small for the split outputs");
  }
  DL_CHECK_ARG(!y || split > 0 || (d->ldy % 8 == 0 && d->ldy >= d->Cout), "conv_igemm: ldy must be >= Cout, multiple of 8");
  DL_CHECK_ARG(!y_f32 || (d->ldf % 4 == 0 && d->ldf >= d->Cout), "conv_igemm: ldf must be >= Cout, multiple of 4");
  DL_CHECK_ARG(d->R >= 1 && d->S >= 1 && d->stride_h >= 1 && d->stride_w >= 1 && d->dil_h >= 1 && d->dil_w >= 1 &&
                   d->pad_h >= 0 && d->pad_w >= 0,
               "conv_igemm: bad filter geometry");
  DL_CHECK_ARG((d->S - 1) * d->dil_w <= 255 && (d->R - 1) * d->dil_h <= 255 && d->pad_w <= 127 && d->pad_h <= 127,
               "conv_igemm: filter extent exceeds the TMA im2col offset range");
  const bool lin = d->lin != 0;
  int P, Q;
  if (lin) {
    DL_CHECK_ARG(d->stride_h == 1 && d->stride_w == 1, "conv_igemm: lin mode needs stride 1");
    DL_CHECK_ARG(d->valid_h >= 1 && d->valid_h <= d->H && d->valid_w >= 1 && d->valid_w <= d->W,
                 "conv_igemm: lin mode needs 1 <= valid_h <= H and 1 <= valid_w <= W");
    DL_CHECK_ARG(d->img_rows == 0 && d->img_cols == 0 && d->out_img_rows == 0 && d->out_img_cols == 0,
                 "conv_igemm: lin mode takes its pitches from H and W");
    P = d->H; Q = d->W;
  } else {
    P = (d->H + 2 * d->pad_h - d->dil_h * (d->R - 1) - 1) / d->stride_h + 1;
    Q = (d->W + 2 * d->pad_w - d->dil_w * (d->S - 1) - 1) / d->stride_w + 1;
    DL_CHECK_ARG(P > 0 && Q > 0, "conv_igemm: input smaller than the filter");
    DL_CHECK_ARG((d->out_img_rows == 0) == (d->out_img_cols == 0), "conv_igemm: out_img_rows/cols must come together");
    DL_CHECK_ARG(d->out_img_rows == 0 || (d->out_img_rows >= P && d->out_img_cols >= Q),
                 "conv_igemm: output pitches smaller than the output");
  }
  int st = require_sm100();
  if (st != DL_OK) return st;

  const long long M = (long long)d->N * P * Q;
  DL_CHECK_ARG(M < (1ll << 31) - 256, "conv_igemm: too many output pixels");
  // ---- a handful of rows through a 1x1 filter (the fc heads: one row per utterance) is latency, not throughput:
  // linear_small_kernel spreads it over Cout/4 blocks instead of one or two tensor-core CTAs
  // (only where one "pixel" is one batch item, P Q == 1, and whatever the batch size up to 4096, so that an utterance
  // gets the same summation order alone and inside a batch; Cout <= 2048 keeps wide "layers" -- the dense trial-scoring
  // GEMM of a small list, 3 526 x 3 447 on LomGRID -- on the tensor-core path: 2.9 ms here vs. tens of microseconds)
  if (opt_small_linear() && split == 0 && M <= 4096 && d->Cout <= 2048 && P * Q == 1 && d->R == 1 && d->S == 1 && d->stride_h == 1 && d->stride_w == 1 && d->pad_h == 0 &&
      d->pad_w == 0 && !lin && !residual && d->C % 8 == 0 && d->out_img_rows == 0 &&
      (d->img_rows == 0 || d->img_rows == d->H) && (d->img_cols == 0 || d->img_cols == d->W) &&
      (((uintptr_t)x | (uintptr_t)w_packed) & 15) == 0) {
    LinearSmallParams lp;
    lp.x = static_cast<const uint16_t*>(x); lp.w = static_cast<const uint16_t*>(w_packed);
    lp.M = (int)M; lp.C = d->C; lp.ldx = d->ldx; lp.ldw = (d->C + 63) / 64 * 64; lp.Cout = d->Cout;
    lp.scale = scale; lp.shift = shift; lp.slope = slope;
    lp.y = static_cast<uint16_t*>(y); lp.ldy = d->ldy;
    lp.scale2 = scale2; lp.shift2 = shift2; lp.f32_slope = d->f32_slope; lp.yf = y_f32; lp.ldf = d->ldf;
    const int gx = (d->Cout + kLinCh - 1) / kLinCh;
    cudaStream_t cs = (cudaStream_t)stream;
    if (M <= 32) linear_small_kernel<1><<<dim3(gx, 1), 32 * kLinKs, linear_small_smem_bytes(1), cs>>>(lp);
    else linear_small_kernel<2><<<dim3(gx, (unsigned)((M + 63) / 64)), 32 * kLinKs, linear_small_smem_bytes(2), cs>>>(lp);
    return check_launch("linear_small_kernel");
  }
  IgemmParams p;
  p.M = (int)M; p.P = P; p.Q = Q; p.PQ = P * Q;
  p.Cout = d->Cout;
  p.cchunks = (d->C + 63) / 64;
  p.R = d->R; p.S = d->S;
  p.stride_h = d->stride_h; p.stride_w = d->stride_w;
  p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  p.dil_h = d->dil_h; p.dil_w = d->dil_w;
  p.ldy = d->ldy; p.ldf = d->ldf;
  p.lin = lin ? 1 : 0; p.lin_w = d->W; p.lin_h = d->H; p.valid_w = d->valid_w; p.valid_h = d->valid_h;
  p.dbg = opt_dbg();
  p.out_hp = lin ? 0 : d->out_img_rows; p.out_wp = lin ? 0 : d->out_img_cols;
  p.f32_slope = d->f32_slope;
  p.scale = scale; p.shift = shift; p.slope = slope; p.scale2 = scale2; p.shift2 = shift2;
  p.residual = static_cast<const uint16_t*>(residual);
  p.y = static_cast<uint16_t*>(y);
  p.yf = y_f32;
  p.y2 = static_cast<uint16_t*>(d->y_split);
  p.split_c = split > 0 ? split : 0;
  p.skip_n0 = -1;
  p.half_skip = 0;
  p.num_m_blocks = (int)((M + 127) / 128);
  p.tile_rows = 128; p.pool_rows = 0; p.store_y = 1; p.pool_out = nullptr;
  const int pool = d->avgpool ? P * Q : 0;
  if (pool > 0)
    DL_CHECK_ARG(d->avgpool_out && y && !y_f32 && split == 0 && !lin && d->out_img_rows == 0 && d->ldy == d->Cout &&
                     (reinterpret_cast<uintptr_t>(d->avgpool_out) & 15) == 0,
                 "conv_igemm: avgpool needs avgpool_out (16-byte aligned), a dense bf16 y (scratch), no f32 / split / guarded output");

  const int block_n = d->Cout <= 64 ? 64 : (d->Cout <= 128 ? 128 : 256);
  p.num_n_blocks = (d->Cout + block_n - 1) / block_n;
  const long long Ktot = (long long)d->R * d->S * p.cchunks * 64;

  const long long num_kb = (long long)d->R * d->S * p.cchunks;
  const bool resident = p.num_n_blocks == 1 && num_kb * block_n * 128 <= kResidentBBytes;
  // CTA pairs (cta_group::2) for wide tiles on problems large enough to fill the chip twice over
  const bool pair = opt_pair() && block_n >= 128 && !resident && p.num_m_blocks >= 128;
  p.taps = pair ? igemm_pair_taps(p, block_n) : 1;
  // the declared centre-tap-only channels: whole n blocks of the pair kernel skip the other K blocks (any other kernel
  // multiplies the zero weights, same result)
  if (split > 0 && d->split_center_only && pair && p.taps == 1 && (d->R & 1) && (d->S & 1) && split % block_n == 0)
    p.skip_n0 = split / block_n;
  p.a_rows = 128 + (p.taps - 1) * d->dil_w;
  // declared centre-tap-only upper half of ONE 256-wide tile (layer2's fused entry block): the resident pair kernel
  // runs the two halves as separate N = 128 MMA streams, the upper one on the centre tap's K blocks only
  DL_CHECK_ARG(d->center_only_from == 0 || (d->center_only_from > 0 && d->center_only_from < d->Cout && split == 0),
               "conv_igemm: center_only_from must lie inside (0, Cout) and excludes split_channel");
  const bool want_half = d->center_only_from == 128 && d->Cout == 256 && p.taps == 1 && (d->R & 1) && (d->S & 1);
  const int staged = pair ? igemm_pair_staged(p, block_n, want_half) : 0;
  p.half_skip = staged == 2 ? 1 : 0;
  // K4 in the epilogue: the streaming 256-wide staged variant reads the image means off its staged output tiles
  const bool pool_fused = pool > 0 && pair && staged == 3 && pool <= 128 && d->Cout % 64 == 0 && opt_pool_fuse();
  if (pool_fused) {
    p.tile_rows = (128 / pool) * pool;
    p.num_m_blocks = (int)((M + p.tile_rows - 1) / p.tile_rows);
    p.pool_rows = pool;
    p.pool_out = d->avgpool_out;
    p.store_y = d->avgpool_keep_y ? 1 : 0;
  }

  CUtensorMap mapA, mapB;
  if (lin) {
    st = make_tiled_2d_bf16(&mapA, x, (uint64_t)M, (uint64_t)d->C, (uint64_t)d->ldx, (uint32_t)p.a_rows, 64);
  } else {
    const int img_rows = d->img_rows > 0 ? d->img_rows : d->H;
    const int img_cols = d->img_cols > 0 ? d->img_cols : d->W;
    DL_CHECK_ARG(img_rows >= d->H && img_cols >= d->W, "conv_igemm: img_rows < H or img_cols < W");
    st = make_im2col_nhwc_bf16(&mapA, x, d->N, d->H, d->W, d->C, d->ldx, img_rows, img_cols, d->R, d->S, d->stride_h,
                               d->stride_w, d->pad_h, d->pad_w, d->dil_h, d->dil_w, 64, 128);
  }
  if (st != DL_OK) return st;
  // weight boxes: a CTA of a pair loads its half of the n block; in half_skip mode as two quarter boxes (64 conv1 rows
  // + 64 skip rows), so that an N = 128 pair MMA finds "its" 64 rows at the same offset in both CTAs
  st = make_tiled_2d_bf16(&mapB, w_packed, (uint64_t)d->Cout, (uint64_t)Ktot, (uint64_t)Ktot,
                          (uint32_t)(p.half_skip ? block_n / 4 : (pair ? block_n / 2 : block_n)), 64);
  if (st != DL_OK) return st;

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pair) {
    CUtensorMap mapY, mapY2;
    if (staged) {       // 64-channel x 128-row boxes of the dense (M, ldy) output; rows past M / channels past Cout are clipped by the TMA unit
      st = make_tiled_2d_bf16(&mapY, y, (uint64_t)M, (uint64_t)(split > 0 ? split : d->Cout), (uint64_t)d->ldy, 128, 64);
      if (st == DL_OK && split > 0)
        st = make_tiled_2d_bf16(&mapY2, d->y_split, (uint64_t)M, (uint64_t)(d->Cout - split), (uint64_t)d->ldy, 128, 64);
      if (st != DL_OK) return st;
    }
    st = launch_igemm_pair(mapA, mapB, p, block_n, s, staged ? &mapY : nullptr, (staged && split > 0) ? &mapY2 : nullptr);
    if (st != DL_OK || pool == 0 || pool_fused) return st;
    return avgpool_unfused(y, d->N, pool, d->Cout, d->avgpool_out, stream);
  }
  switch (block_n) {
    case 64: st = resident ? launch_igemm<64, true>(mapA, mapB, p, s) : launch_igemm<64, false>(mapA, mapB, p, s); break;
    case 128: st = resident ? launch_igemm<128, true>(mapA, mapB, p, s) : launch_igemm<128, false>(mapA, mapB, p, s); break;
    default: st = resident ? launch_igemm<256, true>(mapA, mapB, p, s) : launch_igemm<256, false>(mapA, mapB, p, s); break;
  }
  if (st != DL_OK || pool == 0) return st;
  return avgpool_unfused(y, d->N, pool, d->Cout, d->avgpool_out, stream);
}
