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

template <int BLOCK_N, bool kResB, int TAPS>
struct Igemm2Cfg {
  static_assert(TAPS == 1 || !kResB, "tap sharing streams its weights");
  static constexpr int A_BYTES = (TAPS > 1 ? kTapBoxRows : 128) * 64 * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * 64 * 2;       // this CTA's half of one weight K block
  static constexpr int STAGE_BYTES = kResB ? A_BYTES : A_BYTES + TAPS * B_BYTES;
  static constexpr int BRES_BYTES = kResB ? kPairResidentBBytes : 0;
  static constexpr int PSTRIDE = (kResB || TAPS > 1) ? kPairResidentMaxCout : kMaxCout;   // staged epilogue parameters
  static constexpr int STAGES =
      kResB ? 4 : (TAPS == 1 ? (BLOCK_N == 256 ? 6 : 8) : (225 * 1024 - 3 * PSTRIDE * 4) / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = (2 * STAGES + 5) * 8 + 16;
  static constexpr int PARAM_BYTES = 3 * PSTRIDE * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BRES_BYTES + PARAM_BYTES + BAR_BYTES + 1024;
  static constexpr int THREADS = 320;
  static_assert(STAGES >= 2, "pipeline depth");
};

template <int BLOCK_N, bool kResB, int TAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
igemm2_conv_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const IgemmParams p) {
  using Cfg = Igemm2Cfg<BLOCK_N, kResB, TAPS>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bres = smem + STAGES * Cfg::STAGE_BYTES;            // resident weight half (kResB only)
  float* prm = reinterpret_cast<float*>(bres + Cfg::BRES_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bres + Cfg::BRES_BYTES + Cfg::PARAM_BYTES);
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
        if (rank == 0) mbar_expect_tx_cluster(bb, 2u * (uint32_t)num_kb * Cfg::B_BYTES);
        else mbar_arrive_cluster(bb);
        if (BLOCK_N == 256 && p.half_skip) {
          // this CTA's B block = [conv1 rows 64 rank .. +63 | skip rows 128 + 64 rank .. +63] (two 64-row boxes)
          for (int kb = 0; kb < num_kb; ++kb) {
            tma2_load_2d(bres + kb * Cfg::B_BYTES, &mapB, bb, kb * 64, (int)rank * 64);
            tma2_load_2d(bres + kb * Cfg::B_BYTES + Cfg::B_BYTES / 2, &mapB, bb, kb * 64, 128 + (int)rank * 64);
          }
        } else {
          for (int kb = 0; kb < num_kb; ++kb)
            tma2_load_2d(bres + kb * Cfg::B_BYTES, &mapB, bb, kb * 64, (int)rank * (BLOCK_N / 2));
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < total; t += num_pairs) {
        int mp, n_blk;
        tile_of(t, mp, n_blk);
        const int m0 = (2 * mp + (int)rank) * 128;
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
        if (kResB && BLOCK_N == 256 && p.half_skip) {
          // Two N = 128 streams into one 256-column accumulator: columns [0,128) = conv1 over every K block (B rows
          // 0..63 of each CTA's block), columns [128,256) = the skip conv over the centre tap's K blocks only (B rows
          // 64..127).  36 + 4 half-width MMAs per tile instead of 36 full-width ones; bit-identical (the products
          // left out are exact zeros).
          constexpr uint32_t idesc_h = umma_idesc_bf16(256, 128);
          const int ckb0 = ((p.R >> 1) * p.S + (p.S >> 1)) * p.cchunks;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smem + stage * Cfg::STAGE_BYTES));
            const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(bres) + kb * Cfg::B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc_h, (kb | k) != 0 ? 1u : 0u);
            if (kb >= ckb0 && kb < ckb0 + p.cchunks) {
              const uint64_t sdesc = bdesc + (uint64_t)((Cfg::B_BYTES / 2) >> 4);
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
    for (int t = pair; t < total; t += num_pairs) {
      int mp, n_blk;
      tile_of(t, mp, n_blk);
      long long row = (long long)(2 * mp + (int)rank) * 128 + quarter * 32 + lane;
      const bool row_ok = igemm_map_row(p, row) && !(p.dbg & 2);
      const int cbase = n_blk * BLOCK_N;
      uint4 res[4];
      igemm_prefetch_residual(p, row, row_ok, cbase, chunk0, has_res, res);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (!(p.dbg & 4))
      igemm_epilogue_tile<BLOCK_N>(p, prm, Cfg::PSTRIDE, tmem_base + acc * BLOCK_N, row, row_ok, cbase, quarter, chunk0, has_res, fast,
                                   res);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));   // the leader's MMA thread waits
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody leaves (or frees TMEM) while the peer may still touch this CTA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BLOCK_N, bool kResB, int TAPS>
static int launch_igemm2(const CUtensorMap& mapA, const CUtensorMap& mapB, const IgemmParams& p, cudaStream_t stream) {
  using Cfg = Igemm2Cfg<BLOCK_N, kResB, TAPS>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "shared-memory budget");
  static PerDevice<bool> configured_dev;
  bool* configured = configured_dev.slot();
  if (!configured) return fail(DL_ERR_CUDA, "igemm2: no current device");
  if (!*configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm2_conv_kernel<BLOCK_N, kResB, TAPS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "igemm2 smem attribute: %s", cudaGetErrorString(e));
    *configured = true;
  }
  const int super_tiles = ((p.num_m_blocks + 1) / 2) * p.num_n_blocks;
  int pairs = device_sm_count() / 2;
  if (pairs <= 0) pairs = 74;
  if (super_tiles < pairs) pairs = super_tiles;
  if ((p.dbg & 64) && pairs > 1) pairs /= 2;          // measurement aid: half the SMs (per-SM vs chip-wide ingest)
  igemm2_conv_kernel<BLOCK_N, kResB, TAPS><<<2 * pairs, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(mapA, mapB, p);
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

bool igemm_pair_resident(const IgemmParams& p, int block_n) {
  const long long num_kb = (long long)p.R * p.S * p.cchunks;
  return opt_pair_resident() && p.taps == 1 && p.num_n_blocks == 1 && num_kb * (block_n / 2) * 128 <= kPairResidentBBytes &&
         p.Cout <= kPairResidentMaxCout;
}

// Called by dl_conv_igemm_bf16 (igemm_conv.cu) for wide, large-M problems.
int launch_igemm_pair(const CUtensorMap& mapA, const CUtensorMap& mapB, const IgemmParams& p, int block_n,
                      cudaStream_t stream) {
  if (p.taps == 3)
    return block_n == 128 ? launch_igemm2<128, false, 3>(mapA, mapB, p, stream) : launch_igemm2<256, false, 3>(mapA, mapB, p, stream);
  if (p.taps == 5) return launch_igemm2<256, false, 5>(mapA, mapB, p, stream);
  const bool res = igemm_pair_resident(p, block_n);
  if (block_n == 128) return res ? launch_igemm2<128, true, 1>(mapA, mapB, p, stream) : launch_igemm2<128, false, 1>(mapA, mapB, p, stream);
  return res ? launch_igemm2<256, true, 1>(mapA, mapB, p, stream) : launch_igemm2<256, false, 1>(mapA, mapB, p, stream);
}

}  // namespace dl
