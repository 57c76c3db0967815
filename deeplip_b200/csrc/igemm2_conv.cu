// Implicit-GEMM convolution on CTA pairs (tcgen05 cta_group::2): two SMs of one TPC compute a 256-row x BLOCK_N
// tile together.  Each CTA TMA-im2col-loads its own 128 output pixels of operand A and only HALF of the weight
// tile (operand B); the leader CTA issues one M=256 tcgen05.mma that reads A and the B halves from both CTAs'
// shared memory and writes each CTA's 128 accumulator rows into its own tensor memory.
//
// Why: measured on B200 the tensor pipe's shared-memory operand fetch sustains ~60 B/cycle/SM; a single-CTA
// 128 x 256 x 16 MMA needs 12 KB (A 4 KB + B 8 KB) -> ~190 cycles against a 128-cycle math floor (67 %), and the
// 128 x 128 tile only reaches 50 %.  Halving B per SM brings N=256 to the math floor and N=128 to 67 %.
//
// Same TMA im2col producer / TMEM double buffering / fused epilogue as igemm_conv.cu; barriers of the pair:
//   full[s]   (leader)      : 2 arrivals (each CTA's producer) + all TMA bytes of both CTAs
//   empty[s]  (both)        : tcgen05.commit multicast
//   tfull[a]  (both)        : tcgen05.commit multicast
//   tempty[a] (leader)      : 16 arrivals (one per epilogue warp of both CTAs, the peer's via the cluster window)
#include <cuda_runtime.h>
#include "igemm_common.cuh"

namespace dl {

constexpr int kPairResidentBBytes = 144 * 1024;   // this CTA's half of the whole weight matrix, kept in smem
constexpr int kPairResidentMaxCout = 512;

// kResB: single N block and the CTA's weight half (all K blocks) fits in shared memory -> loaded once per CTA,
// the stages carry operand A only.
// TAPS > 1 (guarded-linear operand A only): one pipeline stage holds ONE A box of 128 + (TAPS-1)*dil_w rows and the
// weight blocks of all TAPS horizontal taps; tap s multiplies the view of the box that starts s*dil_w rows in
// (the 128B swizzle is a function of the absolute shared-memory address, so a view shifted by whole 128-byte rows
// is a valid operand).  A's share of the per-SM ingest (measured ceiling ~50 B/cycle) drops TAPS-fold.
constexpr int kTapBoxRows = 144;                  // A box capacity: 128 + (TAPS-1)*dil_w <= 144

// kStg: output tiles leave through kStg 16 KB shared-memory staging buffers (128 rows x 64 channels, 128B swizzle)
// and one TMA store per 64-channel slab -- full 128-byte lines, asynchronous -- instead of per-lane 16-byte stores at a
// row-pitch stride (measured on the layer2 convs: the direct stores cost 26 of 180 us, on the fused entry 44 of 152).
// Needs a dense row-major y (no guarded layouts) and no f32 side output; a split output brings a second tensor map.
// Short-K layers gain most: on the E-TDNN k=1 layers (K = 512) the per-lane stores were 3.7 of 15.9 us at B = 64 and
// 19 of 48 us at B = 256 (tools/tdnn_fixed.py).
// kHalf (BLOCK_N = 256, resident): the n block's upper half is centre-tap-only (IgemmParams::half_skip): the weights
// are kept compact (64 conv1 rows per K block + 64 skip rows for the centre tap's K blocks only: 80 KB instead of
// 144 KB), which pays for six A stages (the kernel is load-latency bound with four) and two staging buffers.
constexpr int kStgBytes = 128 * 128;
constexpr int kHalfMaxBlocks = 10;                // conv1 K blocks + skip K blocks of the compact layout

template <int BLOCK_N, bool kResB, int TAPS, int kStg = 0, bool kHalf = false>
struct Igemm2Cfg {
  static_assert(TAPS == 1 || !kResB, "tap sharing streams its weights");
  static_assert(kStg == 0 || TAPS == 1, "the staged epilogue is not built for the tap-sharing variants");
  static_assert(kStg == 0 || kResB || (BLOCK_N == 256 && kStg == 2), "streaming-weights staged variant: 256-wide, two buffers");
  static_assert(!kHalf || (kResB && BLOCK_N == 256 && kStg > 0), "half-skip mode: resident 256-wide tile");
  static constexpr int A_BYTES = (TAPS > 1 ? kTapBoxRows : 128) * 64 * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * 64 * 2;       // this CTA's half of one weight K block
  static constexpr int STAGE_BYTES = kResB ? A_BYTES : A_BYTES + TAPS * B_BYTES;
  static constexpr int BRES_BYTES = kResB ? (kHalf ? kHalfMaxBlocks * (B_BYTES / 2) : kPairResidentBBytes) : 0;
  // per-channel epilogue parameters staged in shared memory: one n block (<= BLOCK_N channels) in the kStg variants
  static constexpr int PSTRIDE = (kStg > 0 && kResB) ? BLOCK_N : ((kResB || TAPS > 1) ? kPairResidentMaxCout : kMaxCout);
  // streaming weights + staging: five 32 KB stages instead of six make room for the two 16 KB staging buffers
  static constexpr int STAGES =
      kResB ? (kHalf ? 6 : 4)
            : (TAPS == 1 ? (BLOCK_N == 256 ? (kStg > 0 ? 5 : 6) : 8) : (225 * 1024 - 3 * PSTRIDE * 4) / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = (2 * STAGES + 5) * 8 + 16;
  static constexpr int PARAM_BYTES = 3 * PSTRIDE * 4;
  static constexpr int STG_BYTES = kStg * kStgBytes;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BRES_BYTES + STG_BYTES + PARAM_BYTES + BAR_BYTES + 1024;
  static constexpr int THREADS = 320;
  static_assert(STAGES >= 2, "pipeline depth");
};

template <int BLOCK_N, bool kResB, int TAPS, int kStg, bool kHalf>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
igemm2_conv_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapY2,
                   const IgemmParams p) {
  using Cfg = Igemm2Cfg<BLOCK_N, kResB, TAPS, kStg, kHalf>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bres = smem + STAGES * Cfg::STAGE_BYTES;            // resident weight half (kResB only)
  uint8_t* stg = bres + Cfg::BRES_BYTES;                       // kStg output staging buffers (1024-byte aligned)
  float* prm = reinterpret_cast<float*>(stg + Cfg::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + Cfg::STG_BYTES + Cfg::PARAM_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* bfull = bars + 2 * STAGES + 4;                     // leader: both weight halves landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = (int)cluster_id_x();
  const int num_pairs = (int)num_clusters_x();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (kStg > 0) {
      tma_prefetch_desc(&mapY);
      if (p.split_c > 0) tma_prefetch_desc(&mapY2);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 2);
        mbar_init(&empty[s], 1);
      }
      mbar_init(&tfull[0], 1);
      mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 16);
      mbar_init(&tempty[1], 16);
      mbar_init(bfull, 2);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  }
  if (p.y != nullptr) {
    for (int c = threadIdx.x; c < p.Cout; c += Cfg::THREADS) {
      prm[c] = p.scale[c];
      prm[Cfg::PSTRIDE + c] = p.shift[c];
      prm[2 * Cfg::PSTRIDE + c] = p.slope[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // barriers of both CTAs initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // super tiles: (pair of consecutive m blocks) x n block
  const int num_m_pairs = (p.num_m_blocks + 1) >> 1;
  const int total = num_m_pairs * p.num_n_blocks;
  const int num_kb = p.R * p.S * p.cchunks;
  // centre-tap-only n blocks (p.skip_n0 >= 0) cost a ninth of the others: tiles then go n-block-major, so that every
  // pair works through the expensive ones first and the cheap ones after instead of owning one kind (t = pair + i * pairs
  // with an even pair count would give a pair one n block only)
  const bool ctr_mode = !kResB && TAPS == 1 && p.skip_n0 >= 0;
  auto tile_of = [&](int t, int& mp, int& n_blk) {
    if (ctr_mode) { n_blk = t / num_m_pairs; mp = t - n_blk * num_m_pairs; }
    else { mp = t / p.num_n_blocks; n_blk = t - mp * p.num_n_blocks; }
  };

  if (warp == 0) {
    if (elect_one_sync()) {
      if (kResB) {                       // whole weight half of this CTA, credited to the leader's barrier
        const uint32_t bb = mapa_shared(smem_u32(bfull), 0);
        if (kHalf) {
          // compact layout, 64-row boxes: block kb = conv1 rows 64 rank .. +63 of K block kb; blocks num_kb .. =
          // skip rows 128 + 64 rank .. +63 of the centre tap's K blocks.  An N = 128 pair MMA takes "its" 64 rows
          // from the same offset in both CTAs.
          constexpr uint32_t HB = Cfg::B_BYTES / 2;
          const int ckb0 = ((p.R >> 1) * p.S + (p.S >> 1)) * p.cchunks;
          if (rank == 0) mbar_expect_tx_cluster(bb, 2u * (uint32_t)(num_kb + p.cchunks) * HB);
          else mbar_arrive_cluster(bb);
          for (int kb = 0; kb < num_kb; ++kb) tma2_load_2d(bres + kb * HB, &mapB, bb, kb * 64, (int)rank * 64);
          for (int c = 0; c < p.cchunks; ++c)
            tma2_load_2d(bres + (num_kb + c) * HB, &mapB, bb, (ckb0 + c) * 64, 128 + (int)rank * 64);
        } else {
          if (rank == 0) mbar_expect_tx_cluster(bb, 2u * (uint32_t)num_kb * Cfg::B_BYTES);
          else mbar_arrive_cluster(bb);
          for (int kb = 0; kb < num_kb; ++kb)
            tma2_load_2d(bres + kb * Cfg::B_BYTES, &mapB, bb, kb * 64, (int)rank * (BLOCK_N / 2));
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < total; t += num_pairs) {
        int mp, n_blk;
        tile_of(t, mp, n_blk);
        const int m0 = (2 * mp + (int)rank) * p.tile_rows;
        const int img = m0 / p.PQ;
        const int rem = m0 - img * p.PQ;
        const int pp = rem / p.Q;
        const int qq = rem - pp * p.Q;
        const int w0 = qq * p.stride_w - p.pad_w;
        const int h0 = pp * p.stride_h - p.pad_h;
        if (TAPS > 1) {                  // one A box + the weight blocks of all TAPS horizontal taps per stage
          const uint32_t tx = (uint32_t)p.a_rows * 128u + TAPS * Cfg::B_BYTES;
          for (int r = 0; r < p.R; ++r) {
            for (int cc = 0; cc < p.cchunks; ++cc) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              const uint32_t lbar = mapa_shared(smem_u32(&full[stage]), 0);
              if (rank == 0) mbar_expect_tx_cluster(lbar, 2u * tx);
              else mbar_arrive_cluster(lbar);
              tma2_load_2d(sa, &mapA, lbar, cc * 64, m0 + (r * p.dil_h - p.pad_h) * p.lin_w - p.pad_w);
#pragma unroll
              for (int s = 0; s < TAPS; ++s)
                tma2_load_2d(sa + Cfg::A_BYTES + s * Cfg::B_BYTES, &mapB, lbar, ((r * TAPS + s) * p.cchunks + cc) * 64,
                             n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2));
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
        const bool ctr = ctr_mode && n_blk >= p.skip_n0;
        const int r_lo = ctr ? p.R >> 1 : 0, r_hi = ctr ? r_lo + 1 : p.R;
        const int s_lo = ctr ? p.S >> 1 : 0, s_hi = ctr ? s_lo + 1 : p.S;
        for (int r = r_lo; r < r_hi; ++r) {
          for (int s = s_lo; s < s_hi; ++s) {
            int kb = (r * p.S + s) * p.cchunks;
            for (int cc = 0; cc < p.cchunks; ++cc, ++kb) {
              mbar_wait(&empty[stage], phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
              const uint32_t lbar = mapa_shared(smem_u32(&full[stage]), 0);       // the leader's barrier
              if (rank == 0) mbar_expect_tx_cluster(lbar, 2u * Cfg::STAGE_BYTES);
              else mbar_arrive_cluster(lbar);
              if (p.lin)                   // guarded-linear A: the tile's rows shifted by the tap, plain 2D box
                tma2_load_2d(sa, &mapA, lbar, cc * 64, m0 + (r * p.dil_h - p.pad_h) * p.lin_w + s * p.dil_w - p.pad_w);
              else
                tma2_load_im2col_4d(sa, &mapA, lbar, cc * 64, w0, h0, img, (uint16_t)(s * p.dil_w),
                                    (uint16_t)(r * p.dil_h));
              if (!kResB)
                tma2_load_2d(sa + Cfg::A_BYTES, &mapB, lbar, kb * 64, n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2));
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one_sync()) {   // one thread, known to ptxas as such: no per-lane loops around UTCHMMA
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N);
      if (kResB) mbar_wait(bfull, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair; t < total; t += num_pairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BLOCK_N;
        if (TAPS > 1) {
          const int iters = p.R * p.cchunks;
          for (int it = 0; it < iters; ++it) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
            for (int s = 0; s < TAPS; ++s) {
              const uint64_t adesc = umma_desc_sw128_kmajor(sa + (uint32_t)(s * p.dil_w) * 128u);
              const uint64_t bdesc = umma_desc_sw128_kmajor(sa + Cfg::A_BYTES + s * Cfg::B_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma2_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (it | s | k) != 0 ? 1u : 0u);
            }
            umma2_commit_both(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
        if (kHalf) {
          // Two N = 128 streams into one 256-column accumulator: columns [0,128) = conv1 over every K block,
          // columns [128,256) = the skip conv over the centre tap's K blocks only.  36 + 4 half-width MMAs per tile
          // instead of 36 full-width ones; bit-identical (the products left out are exact zeros).
          constexpr uint32_t idesc_h = umma_idesc_bf16(256, 128);
          const int ckb0 = ((p.R >> 1) * p.S + (p.S >> 1)) * p.cchunks;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem + stage * Cfg::STAGE_BYTES));
            const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(bres) + kb * (Cfg::B_BYTES / 2));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc_h, (kb | k) != 0 ? 1u : 0u);
            if (kb >= ckb0 && kb < ckb0 + p.cchunks) {
              const uint64_t sdesc = umma_desc_sw128_kmajor(smem_u32(bres) + (num_kb + kb - ckb0) * (Cfg::B_BYTES / 2));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma2_bf16(d + 128, adesc + 2 * k, sdesc + 2 * k, idesc_h, (kb != ckb0 || k != 0) ? 1u : 0u);
            }
            umma2_commit_both(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
        int mp_, n_blk_;
        tile_of(t, mp_, n_blk_);
        const int nkb = (ctr_mode && n_blk_ >= p.skip_n0) ? p.cchunks : num_kb;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128_kmajor(sa);
          const uint64_t bdesc = umma_desc_sw128_kmajor(kResB ? smem_u32(bres) + kb * Cfg::B_BYTES : sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < ((p.dbg & 8) ? 1 : 4); ++k) umma2_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_both(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        }
        }
        umma2_commit_both(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    const bool has_res = p.residual != nullptr && !(p.dbg & 1);
    const bool fast = p.y != nullptr && p.yf == nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    int sbuf = 0;                                  // kStg: staging buffer of the next slab
    const bool issuer = warp == 2 && lane == 0;    // kStg: the thread that owns this CTA's TMA stores
    for (int t = pair; t < total; t += num_pairs) {
      int mp, n_blk;
      tile_of(t, mp, n_blk);
      const int tile_row0 = (2 * mp + (int)rank) * p.tile_rows;
      long long row = (long long)tile_row0 + quarter * 32 + lane;
      const bool row_ok = igemm_map_row(p, row) && !(p.dbg & 2);
      const int cbase = n_blk * BLOCK_N;
      uint4 res[4];
      igemm_prefetch_residual(p, row, row_ok, cbase, chunk0, has_res, res);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (!(p.dbg & 4)) {
        if constexpr (kStg > 0)
          igemm_epilogue_tile_staged<BLOCK_N, kStg>(p, prm, Cfg::PSTRIDE, tmem_base + acc * BLOCK_N, row, row_ok, tile_row0, cbase,
                                                    quarter, chunk0, has_res, res, stg, sbuf,
                                                    (p.split_c > 0 && cbase >= p.split_c) ? &mapY2 : &mapY,
                                                    (p.split_c > 0 && cbase >= p.split_c) ? cbase - p.split_c : cbase, issuer);
        else
          igemm_epilogue_tile<BLOCK_N>(p, prm, Cfg::PSTRIDE, tmem_base + acc * BLOCK_N, row, row_ok, cbase, quarter, chunk0,
                                       has_res, fast, res);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));   // the leader's MMA thread waits
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  if (kStg > 0 && warp == 2 && lane == 0) bulk_wait_group0();     // this CTA's outstanding output stores
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BLOCK_N, bool kResB, int TAPS, int kStg = 0, bool kHalf = false>
static int launch_igemm2(const CUtensorMap& mapA, const CUtensorMap& mapB, const IgemmParams& p, cudaStream_t stream,
                         const CUtensorMap* mapY = nullptr, const CUtensorMap* mapY2 = nullptr) {
  using Cfg = Igemm2Cfg<BLOCK_N, kResB, TAPS, kStg, kHalf>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared-memory budget");
  static PerDevice<bool> configured_dev;
  bool* configured = configured_dev.slot();
  if (!configured) return fail(DL_ERR_CUDA, "igemm2: no current device");
  if (!*configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm2_conv_kernel<BLOCK_N, kResB, TAPS, kStg, kHalf>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "igemm2 smem attribute: %s", cudaGetErrorString(e));
    *configured = true;
  }
  const int super_tiles = ((p.num_m_blocks + 1) / 2) * p.num_n_blocks;
  int pairs = device_sm_count() / 2;
  if (pairs <= 0) pairs = 74;
  if (super_tiles < pairs) pairs = super_tiles;
  if ((p.dbg & 64) && pairs > 1) pairs /= 2;          // measurement aid: half the SMs (per-SM vs chip-wide ingest)
  igemm2_conv_kernel<BLOCK_N, kResB, TAPS, kStg, kHalf><<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(
      mapA, mapB, mapY ? *mapY : mapA, mapY2 ? *mapY2 : mapA, p);
  return check_launch("igemm2_conv_kernel");
}

// Horizontal taps one pipeline stage of the pair kernel shares an A box across (1 = none); the caller sizes
// operand A's TMA box to 128 + (taps - 1) * dil_w rows.
int igemm_pair_taps(const IgemmParams& p, int block_n) {
  if (!p.lin || !opt_tap_share() || (p.S - 1) * p.dil_w > kTapBoxRows - 128 || p.Cout > kPairResidentMaxCout) return 1;
  if (p.S == 3) return 3;
  if (p.S == 5 && block_n == 256) return 5;
  return 1;
}

bool igemm_pair_resident(const IgemmParams& p, int block_n);

// Which staged variant (if any) launch_igemm_pair would pick: 0 none, 1 = 128-wide resident tile, 2 = half-skip entry tile
// (compact weights: at most kHalfMaxBlocks 64-row blocks), 3 = 256-wide tiles with streaming weights.
int igemm_pair_staged(const IgemmParams& p, int block_n, bool want_half) {
  if (!opt_staged_epilogue() || p.taps != 1 || p.y == nullptr || p.yf != nullptr || p.lin || p.out_wp > 0 ||
      (reinterpret_cast<uintptr_t>(p.y) & 15) != 0 || (p.split_c > 0 && (reinterpret_cast<uintptr_t>(p.y2) & 15) != 0))
    return 0;
  if (!igemm_pair_resident(p, block_n)) return block_n == 256 ? 3 : 0;      // streaming weights, 256-wide tiles
  if (p.split_c > 0) return 0;
  if (block_n == 256 && want_half && p.R * p.S * p.cchunks + p.cchunks <= kHalfMaxBlocks) return 2;
  if (block_n == 128) return 1;
  return 0;
}

bool igemm_pair_resident(const IgemmParams& p, int block_n) {
  const long long num_kb = (long long)p.R * p.S * p.cchunks;
  return opt_pair_resident() && p.taps == 1 && p.num_n_blocks == 1 && num_kb * (block_n / 2) * 128 <= kPairResidentBBytes &&
         p.Cout <= kPairResidentMaxCout;
}

// Called by dl_conv_igemm_bf16 (igemm_conv.cu) for wide, large-M problems.
int launch_igemm_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const IgemmParams& p, int block_n,
                      cudaStream_t stream, const CUtensorMap* mapY, const CUtensorMap* mapY2) {
  if (p.taps == 3)
    return block_n == 128 ? launch_igemm2<128, false, 3>(mapA, mapB, p, stream) : launch_igemm2<256, false, 3>(mapA, mapB, p, stream);
  if (p.taps == 5) return launch_igemm2<256, false, 5>(mapA, mapB, p, stream);
  const bool res = igemm_pair_resident(p, block_n);
  // mapY != nullptr: the caller found the output eligible for the staged (TMA store) epilogue, see igemm_pair_staged
  if (res && mapY != nullptr) {
    if (block_n == 256 && p.half_skip) return launch_igemm2<256, true, 1, 2, true>(mapA, mapB, p, stream, mapY);
    // 128-wide resident tile: four A stages + ONE staging buffer is all that fits next to the 144 KB of weights (three
    // stages + two buffers measured 218 vs 171 us on the layer2 convs: the im2col pipeline needs its depth)
    // (a streaming-weights 128-wide variant with 7 stages + two staging buffers measured level: 175.6 vs 174.4 us)
    if (block_n == 128) return launch_igemm2<128, true, 1, 1, false>(mapA, mapB, p, stream, mapY);
  }
  if (!res && mapY != nullptr && block_n == 256) return launch_igemm2<256, false, 1, 2, false>(mapA, mapB, p, stream, mapY, mapY2);
  if (block_n == 128) return res ? launch_igemm2<128, true, 1>(mapA, mapB, p, stream) : launch_igemm2<128, false, 1>(mapA, mapB, p, stream);
  return res ? launch_igemm2<256, true, 1>(mapA, mapB, p, stream) : launch_igemm2<256, false, 1>(mapA, mapB, p, stream);
}

}  // namespace dl
