// K2, second generation: the video stem with CHANNELS on the accumulator's lanes and PIXELS on its columns.
//   Conv3d(1->64, k(5,7,7), s(1,2,2), p(2,3,3)) + BatchNorm3d + PReLU + MaxPool3d((1,3,3),(1,2,2),(0,1,1))
//   (models/video_models/model.py:81-85), emitting per-frame NHWC maps like stem_conv3d_kernel.
//
// Why transposed.  The first-generation kernel (stem_conv3d.cu: pixels on lanes) is paced by its epilogue and by the
// shared-memory port: BN parameters are broadcast loads, every conv value goes through a shared-memory ring and is read
// back nine times by the pooling pass, and 8 builder warps assemble im2col rows pixel by pixel.
// Here D^T = W * im2col^T:
//   * operand A = the WEIGHTS, [128 x 64] per input frame: rows 0..63 = W[kt] for output frame t0, rows 64..127 =
//     W[kt-1] for frame t0+1 (a frame pair shares 4 of its 6 input frames), a 128-row window of the resident stack
//     [0, W4, W3, W2, W1, W0, 0] (K-major, 128B swizzle, 56 KB);
//   * operand B = the input WINDOWS, N = 4 conv rows x Wo pixels = 176, K = 16 per MMA = window rows (kh, kh+1) x 8
//     columns.  B is never built pixel by pixel: builder warps "unfold" each input row ONCE into Wo chunks of 16 bytes
//     (chunk ox = input columns 2ox-3 .. 2ox+4), odd input rows in one plane, even rows in the other.  In that layout
//     the operand for window-row pair j is a plain no-swizzle K-major VIEW: pixel n = oy*Wo + ox sits at byte
//     j*Wo*16 + n*16 of the odd plane (kh = 2j) and at the same offset of the even plane (kh = 2j+1), i.e.
//     SBO = 128 B between 8-pixel groups, LBO = plane size between the two K halves.  616 chunks per stage instead of
//     176 x 7 im2col chunks, and no tensor-memory stores;
//   * the accumulator row of a thread is ONE channel of ONE frame over 176 pixels: BN + PReLU use three per-thread
//     scalars, the 3x3/2 max-pool runs in registers (the last conv row of a tile is carried to the next tile) -- on the
//     RAW accumulators where the channel's BN + PReLU is monotone (slope >= 0), so that BN + PReLU touch 22 pooled
//     values per thread and tile instead of 96 -- and the pooled rows leave through a 11 KB staging buffer and two bulk
//     stores of 5.6 KB (two whole NHWC rows per frame).
//
// Work unit = two consecutive output frames of a clip; tile = 4 conv rows (2 pooled rows); 6 pipeline stages per tile
// (input frames t0-2 .. t0+3), 3 MMAs of 128 x 176 x 16 per stage for window rows 0..5 + one MMA per stage PAIR for
// window row 6 (its two K halves are the same place in two consecutive stages): 21 per tile.  The units x tiles
// sequence is cut into equal contiguous ranges, one per CTA.  Shapes: W = 88 (the corpus crop; Wo = 44), H % 8 == 0;
// other shapes run on the first-generation kernel.
#pragma once
#include "dl_host.cuh"
#include "dl_ptx.cuh"

namespace dl {

constexpr int kS2Wo = 44;                           // conv columns
constexpr int kS2Wp = kS2Wo / 2;                    // pooled columns
constexpr int kS2N = 4 * kS2Wo;                     // pixels per tile (MMA N)
constexpr int kS2RowBytes = kS2Wo * 16;             // one unfolded input row
constexpr int kS2Plane = 7 * kS2RowBytes;           // 7 unfolded rows of one parity
constexpr int kS2UStage = 2 * kS2Plane;             // 9 856 B
constexpr int kS2StripRows = 14;                    // input rows 8j-3 .. 8j+10 of tile j
constexpr int kS2Pitch = 2 * kS2Wo + 8;             // pre-pass row pitch in elements (column c = ix + 3)
constexpr int kS2StripBytes = kS2StripRows * kS2Pitch * 2;
constexpr int kS2TileStrip = 6 * kS2StripBytes;     // one TMA box per (frame pair, tile): 6 input frames x 14 rows
constexpr int kS2StripSlots = 4;                    // tiles of strips in flight (the loads are latency-bound)
constexpr int kS2Chunks = kS2StripRows * kS2Wo;     // 616 chunks of 16 B per stage
constexpr int kS2SlotsPerRow = 48;                  // builder work items per strip row (44 chunks + 4 idle): a warp never straddles banks
constexpr int kS2PairBytes = 128 * 128;             // [128 x 64] block holding the three kh = 6 weight pairs (K slices 0..2)
constexpr int kS2Stages = 6;
constexpr int kS2StackBytes = 7 * 8192;             // [0, W4, W3, W2, W1, W0, 0], 64 rows x 128 B each
constexpr int kS2StgBytes = 2 * 2 * kS2Wp * 128;    // [frame][pooled row][px][64 ch] bf16
constexpr int kS2BuilderWarps = 6;                  // two groups of three alternate stages: 96 threads = two strip rows of 48 slots
constexpr int kS2Threads = (8 + 2 + kS2BuilderWarps) * 32;
constexpr int kS2EpiThreads = 256;
constexpr int kS2TmemCols = 512;                    // two accumulators of 176 columns at 0 and 256
constexpr size_t kS2Smem = 1024 + kS2StackBytes + kS2PairBytes + kS2Stages * kS2UStage + kS2StripSlots * kS2TileStrip + 2 * kS2StgBytes + 512;

struct Stem2Params {
  int B, T;
  int tiles;            // Ho / 4
  int pairs_per_clip, units;
  int out_img_rows;     // rows per output frame (>= Hp)
  const float* scale;
  const float* shift;
  const float* slope;
  uint16_t* y;
  const uint16_t* w;    // packed weights (64, 320): K = kt*64 + kh*8 + kw
  int dbg;
};

// No-swizzle K-major operand: core matrices of 8 rows x 16 B (128 B contiguous); `sbo` bytes between 8-row groups,
// `lbo` bytes between the two K halves of one K = 16 step.
__device__ __forceinline__ uint64_t umma_desc_noswizzle_kmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
// shared -> global bulk copy (bulk async-group completion), 16-byte granular
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}

// Work split: the units x tiles sequence (tile-major inside a unit) is cut into gridDim.x contiguous ranges [g0, g1) of
// (nearly) equal length, so that no CTA waits for a last round of whole units (2 432 units on 148 CTAs would be 17
// rounds of 11 tiles against 16.4).  A range that starts inside a unit first recomputes the tile before it (gs = g0 - 1,
// nothing stored): the first pooled row needs that tile's last conv row.
struct Stem2Range { int gs, g0, g1; };
__device__ __forceinline__ Stem2Range stem2_range(const Stem2Params& p) {
  const int total = p.units * p.tiles;
  const int per = total / (int)gridDim.x, rem = total - per * (int)gridDim.x;
  const int b = (int)blockIdx.x;
  Stem2Range r;
  r.g0 = b * per + min(b, rem);
  r.g1 = r.g0 + per + (b < rem ? 1 : 0);
  r.gs = (r.g0 % p.tiles) ? r.g0 - 1 : r.g0;
  return r;
}

// BN + PReLU on 24 consecutive conv columns of one row, then the horizontal 3-max at stride 2.
// HALF 0: columns 0..23 loaded, pooled px 0..10 (px 0 has no left neighbour); HALF 1: columns 20..43, px 11..21.
template <int HALF>
__device__ __forceinline__ void stem2_row(const uint32_t (&v)[24], float sc, float sh, float sl, float (&hp)[11]) {
  float z[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    const float t = fmaf(__uint_as_float(v[k]), sc, sh);
    z[k] = t > 0.f ? t : t * sl;
  }
  if (HALF == 0) {
    hp[0] = fmaxf(z[0], z[1]);
#pragma unroll
    for (int i = 1; i < 11; ++i) hp[i] = fmaxf(fmaxf(z[2 * i - 1], z[2 * i]), z[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 11; ++i) hp[i] = fmaxf(fmaxf(z[2 * i + 1], z[2 * i + 2]), z[2 * i + 3]);
  }
}

// Pool-first form for channels whose BN + PReLU is monotone (slope >= 0): max(f(v)) == f(max v) for scale >= 0 and
// == f(min v) for scale <= 0 -- exactly, f being applied to one of the v either way -- so the 3x3/2 pooling runs on the
// RAW accumulators (FMNMX3 only) and BN + PReLU touch the 22 pooled values of a tile instead of its 96 conv values.
// kMax: horizontal max (else min) of the raw columns, same windows as stem2_row.
template <int HALF, bool kMax>
__device__ __forceinline__ void stem2_row_raw(const uint32_t (&v)[24], float (&hp)[11]) {
  auto op = [](float a, float b) { return kMax ? fmaxf(a, b) : fminf(a, b); };
  float z[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) z[k] = __uint_as_float(v[k]);
  if (HALF == 0) {
    hp[0] = op(z[0], z[1]);
#pragma unroll
    for (int i = 1; i < 11; ++i) hp[i] = op(op(z[2 * i - 1], z[2 * i]), z[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < 11; ++i) hp[i] = op(op(z[2 * i + 1], z[2 * i + 2]), z[2 * i + 3]);
  }
}

// Epilogue modes (one per warp, from its 32 channels' parameters): 0 = general (BN + PReLU on every conv value, then
// the max-pool: any parameters), 1 = pool-first, all scales >= 0 (max of the raw values), 2 = pool-first with mixed
// signs of the scale (max AND min of the raw values, the thread keeps the one its channel needs).
// Epilogue of one warp over all units of this CTA.  q = TMEM lane quarter: lanes 0..63 = frame t0, 64..127 = frame t0+1.
template <int HALF, int MODE>
__device__ __forceinline__ void stem2_epilogue(const Stem2Params& p, uint32_t tmem_base, uint64_t* tfull,
                                               uint64_t* tempty, uint8_t* stg, int q, int lane) {
  const int g = q >> 1;
  const int ch = (q & 1) * 32 + lane;
  const float sc = __ldg(p.scale + ch), sh = __ldg(p.shift + ch), sl = __ldg(p.slope + ch);
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (HALF ? 20u : 0u);
  const bool issuer = threadIdx.x == 0;
  const int px0 = HALF ? 11 : 0;
  int acc = 0, buf = 0;
  uint32_t acc_phase = 0;
  const Stem2Range rg = stem2_range(p);
  int unit = rg.gs / p.tiles, tile = rg.gs - unit * p.tiles;
  int nfr = 0;
  uint16_t* yframe0 = nullptr;
  float carry[11];
  float carry_mn[MODE == 2 ? 11 : 1];          // MODE 2: the min-pooled raw row next to the max-pooled one
  const bool use_min = sc < 0.f;               // MODE 2: this channel's BN turns the order around
  {
    for (int gi = rg.gs; gi < rg.g1; ++gi) {
      if (tile == 0 || gi == rg.gs) {
        const int fb = unit / p.pairs_per_clip;
        const int t0 = 2 * (unit - fb * p.pairs_per_clip);
        nfr = min(2, p.T - t0);
        yframe0 = p.y + ((size_t)fb * p.T + t0) * p.out_img_rows * kS2Wp * 64;
#pragma unroll
        for (int i = 0; i < 11; ++i) {                       // conv row -1 does not exist (a warm-up tile sets it for real)
          carry[i] = -INFINITY;
          if (MODE == 2) carry_mn[i] = INFINITY;
        }
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t a0 = lane_addr + acc * 256;
      float pa[11], pb[11];
      {
        uint32_t v0[24], v1[24], v2[24], v3[24];
        tmem_ld_32x32_x16(a0, v0);
        tmem_ld_32x32_x8(a0 + 16, v0 + 16);
        tmem_ld_32x32_x16(a0 + kS2Wo, v1);
        tmem_ld_32x32_x8(a0 + kS2Wo + 16, v1 + 16);
        tmem_ld_32x32_x16(a0 + 2 * kS2Wo, v2);
        tmem_ld_32x32_x8(a0 + 2 * kS2Wo + 16, v2 + 16);
        tmem_ld_32x32_x16(a0 + 3 * kS2Wo, v3);
        tmem_ld_32x32_x8(a0 + 3 * kS2Wo + 16, v3 + 16);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty[acc]);                    // the accumulator is in registers: the next tile's MMAs may start
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        if (MODE == 0) {
          float h0[11], h1[11];
          if (p.dbg & 64) {                // timing emulation: no BN / PReLU / pooling arithmetic
#pragma unroll
            for (int i = 0; i < 11; ++i) { h0[i] = __uint_as_float(v0[i]); h1[i] = __uint_as_float(v1[i]); }
          } else {
            stem2_row<HALF>(v0, sc, sh, sl, h0);
            stem2_row<HALF>(v1, sc, sh, sl, h1);
          }
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            pa[i] = fmaxf(fmaxf(carry[i], h0[i]), h1[i]);
            pb[i] = h1[i];
          }
          float h2[11], h3[11];
          if (p.dbg & 64) {
#pragma unroll
            for (int i = 0; i < 11; ++i) { h2[i] = __uint_as_float(v2[i]); h3[i] = __uint_as_float(v3[i]); }
          } else {
            stem2_row<HALF>(v2, sc, sh, sl, h2);
            stem2_row<HALF>(v3, sc, sh, sl, h3);
          }
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            pb[i] = fmaxf(fmaxf(pb[i], h2[i]), h3[i]);
            carry[i] = h3[i];
          }
        } else {
          // pool the raw accumulators, then BN + PReLU on the 22 pooled values
          float h0[11], h1[11], h2[11], h3[11];
          stem2_row_raw<HALF, true>(v0, h0);
          stem2_row_raw<HALF, true>(v1, h1);
          stem2_row_raw<HALF, true>(v2, h2);
          stem2_row_raw<HALF, true>(v3, h3);
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            pa[i] = fmaxf(fmaxf(carry[i], h0[i]), h1[i]);
            pb[i] = fmaxf(fmaxf(h1[i], h2[i]), h3[i]);
            carry[i] = h3[i];
          }
          if (MODE == 2) {
            float qa[11], qb[11];
            stem2_row_raw<HALF, false>(v0, h0);
            stem2_row_raw<HALF, false>(v1, h1);
            stem2_row_raw<HALF, false>(v2, h2);
            stem2_row_raw<HALF, false>(v3, h3);
#pragma unroll
            for (int i = 0; i < 11; ++i) {
              qa[i] = fminf(fminf(carry_mn[i], h0[i]), h1[i]);
              qb[i] = fminf(fminf(h1[i], h2[i]), h3[i]);
              carry_mn[i] = h3[i];
              pa[i] = use_min ? qa[i] : pa[i];
              pb[i] = use_min ? qb[i] : pb[i];
            }
          }
#pragma unroll
          for (int i = 0; i < 11; ++i) {
            const float ta = fmaf(pa[i], sc, sh), tb = fmaf(pb[i], sc, sh);
            pa[i] = ta > 0.f ? ta : ta * sl;
            pb[i] = tb > 0.f ? tb : tb * sl;
          }
        }
      }
      // pooled rows 2*tile and 2*tile+1 of frame g, pixels px0 .. px0+10, channel ch -> staging [g][row][px][ch]
      uint8_t* sb = stg + buf * kS2StgBytes + ((g * 2) * kS2Wp + px0) * 128 + ch * 2;
      if (!(p.dbg & 32)) {
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          *reinterpret_cast<__nv_bfloat16*>(sb + i * 128) = __float2bfloat16_rn(pa[i]);
          *reinterpret_cast<__nv_bfloat16*>(sb + (kS2Wp + i) * 128) = __float2bfloat16_rn(pb[i]);
        }
      }
      fence_proxy_async_smem();                       // generic-proxy writes -> visible to the bulk-copy unit
      // the previous tile's stores (other buffer) must have read their buffer before anyone passes this barrier and
      // starts on the tile after this one
      if (issuer) bulk_wait_group_read0();
      named_bar_sync(2, kS2EpiThreads);
      if (issuer && gi >= rg.g0 && !(p.dbg & 2)) {
        const uint8_t* src = stg + buf * kS2StgBytes;
        uint16_t* dst = yframe0 + (size_t)(2 * tile) * kS2Wp * 64;
        bulk_store_s2g(dst, src, 2 * kS2Wp * 128);
        if (nfr == 2) bulk_store_s2g(dst + (size_t)p.out_img_rows * kS2Wp * 64, src + 2 * kS2Wp * 128, 2 * kS2Wp * 128);
        bulk_commit_group();
      }
      buf ^= 1;
      if (++tile == p.tiles) { tile = 0; ++unit; }
    }
  }
  if (issuer) bulk_wait_group0();
}

__global__ void __launch_bounds__(kS2Threads, 1)
stem2_conv3d_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
                    const Stem2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stack = smem;                                          // 7 blocks of [64 cout x 64 K], 128B swizzle
  uint8_t* pairb = stack + kS2StackBytes;                         // kh = 6 weight pairs of stages (0,1), (2,3), (4,5)
  uint8_t* ubuf = pairb + kS2PairBytes;                           // 6 stages x 2 planes x 7 unfolded rows
  uint8_t* strips = ubuf + kS2Stages * kS2UStage;                 // 4 slots x 6 input frames x 14 rows
  uint8_t* stg = strips + kS2StripSlots * kS2TileStrip;           // 2 output staging buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 2 * kS2StgBytes);
  uint64_t* ufull = bars;                        // [6] unfolded stage written (one arrive per builder warp of the group)
  uint64_t* uempty = bars + kS2Stages;           // [6] the stage's MMAs have completed
  uint64_t* sfull = bars + 2 * kS2Stages;        // [4] a tile's strips landed (TMA)
  uint64_t* sempty = bars + 3 * kS2Stages;       // [4] consumed by every builder warp
  uint64_t* tfull = bars + 4 * kS2Stages;        // [2]
  uint64_t* tempty = tfull + 2;                  // [2]
  uint64_t* wbar = tempty + 2;                   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // zero weight blocks at both ends of the stack (the frame of the pair an edge input does not reach)
  for (int i = threadIdx.x; i < 8192 / 16; i += kS2Threads) {
    reinterpret_cast<uint4*>(stack)[i] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(stack + 6 * 8192)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // The 8th window-row slot of a stage is K padding.  Instead of multiplying it, window row kh = 6 of stages 2m and 2m+1
  // share ONE K = 16 MMA (21 instead of 24 MMAs per tile): K slice m of this block holds [W[kt][6][0..7] of stage 2m |
  // of stage 2m+1], kt = stage for rows 0..63 (frame t0), stage - 1 for rows 64..127 (frame t0+1); 128B swizzle.
  for (int i = threadIdx.x; i < 128 * 6; i += kS2Threads) {
    const int r = i / 6, c = i - r * 6;
    const int kt = c - (r >> 6);                          // stage index = c; frame t0+1 sees the previous tap
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (kt >= 0 && kt <= 4) v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)(r & 63) * 320 + kt * 64 + 48));
    *reinterpret_cast<uint4*>(pairb + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < kS2Stages; ++s) {
        mbar_init(&ufull[s], kS2BuilderWarps / 2);
        mbar_init(&uempty[s], 1);
      }
      for (int s = 0; s < kS2StripSlots; ++s) {
        mbar_init(&sfull[s], 1);
        mbar_init(&sempty[s], kS2BuilderWarps);
      }
      mbar_init(&tfull[0], 1);
      mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], kS2EpiThreads);
      mbar_init(&tempty[1], kS2EpiThreads);
      mbar_init(wbar, 1);
      fence_mbar_init();
      tma_prefetch_desc(&mapW);
      tma_prefetch_desc(&mapX);
    }
    __syncwarp();
    tmem_alloc<kS2TmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 10) {
    // =============================================================== builders: unfold input rows into windows
    // chunk (strip row rr, window ox): 16 bytes = input columns 2ox-3 .. 2ox+4 = four aligned words of the strip;
    // consecutive lanes read consecutive words and write consecutive 16-byte chunks; work items are numbered with 48
    // slots per row (4 idle), so the rows a warp straddles start 16 lanes apart: no bank conflicts.
    const int group = (warp - 10) / (kS2BuilderWarps / 2);
    const int gt = (warp - 10 - group * (kS2BuilderWarps / 2)) * 32 + lane;      // 0..95
    static_assert((kS2BuilderWarps / 2) * 32 == 2 * kS2SlotsPerRow, "a group covers two strip rows per round");
    // thread = (row parity `half`, window ox): strip rows half, half+2, ..., half+12 -> unfolded rows 0..6 of plane
    // `half`; every address below is a per-thread constant plus an immediate
    const int half = gt / kS2SlotsPerRow, ox = gt - half * kS2SlotsPerRow;
    const bool ok = ox < kS2Wo;
    const uint32_t src0 = smem_u32(strips) + half * (kS2Pitch * 2) + ox * 4;
    const uint32_t dst0 = smem_u32(ubuf) + half * kS2Plane + ox * 16;
    uint32_t ph = 0, sslot = 0, sph = 0;
    const Stem2Range rg = stem2_range(p);
    {
      for (int gi = rg.gs; gi < rg.g1; ++gi) {
        mbar_wait(&sfull[sslot], sph);
#pragma unroll 1
        for (int st = group; st < kS2Stages; st += 2) {
          const uint32_t src = src0 + sslot * kS2TileStrip + st * kS2StripBytes;
          const uint32_t dst = dst0 + st * kS2UStage;
          uint32_t w[7][4];
#pragma unroll
          for (int r = 0; r < 7; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              w[r][k] = 0u;
              if (ok && !(p.dbg & 16))
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[r][k]) : "r"(src + r * (2 * kS2Pitch * 2) + k * 4));
            }
          }
          if (p.dbg & 256) {             // timing emulation: the strip reads happen, zeros are stored
#pragma unroll
            for (int r = 0; r < 7; ++r) {
              asm volatile("" ::"r"(w[r][0]), "r"(w[r][1]), "r"(w[r][2]), "r"(w[r][3]));
              w[r][0] = w[r][1] = w[r][2] = w[r][3] = 0u;
            }
          }
          mbar_wait(&uempty[st], ph ^ 1);                 // the MMAs that read this stage have completed
          if (ok) {
#pragma unroll
            for (int r = 0; r < 7; ++r)
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + r * kS2RowBytes), "r"(w[r][0]),
                           "r"(w[r][1]), "r"(w[r][2]), "r"(w[r][3])
                           : "memory");
          }
          fence_proxy_async_smem();                       // generic-proxy writes -> visible to the tensor core's reads
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&ufull[st]);
            if (st + 2 >= kS2Stages) mbar_arrive(&sempty[sslot]);   // this warp's share of the tile's strips is unfolded
          }
        }
        ph ^= 1;
        if (++sslot == kS2StripSlots) { sslot = 0; sph ^= 1; }
      }
    }
  } else if (warp == 9) {
    // =============================================================== strip producer: one thread, TMA only
    // one box per (frame pair, tile): rows 8*tile .. 8*tile+13 of the pre-pass frames (row = iy + 3) of input frames
    // t0-2 .. t0+3; a frame index outside [0,T) is out of bounds in the T dimension -> zero fill = Conv3d's temporal
    // padding.  Four tiles of strips are in flight: a 16 KB box takes longer to arrive than a tile takes to compute.
    if (elect_one_sync()) {
      uint32_t slot = 0, ph = 0;
      const Stem2Range rg = stem2_range(p);
      int unit = rg.gs / p.tiles, tile = rg.gs - unit * p.tiles;
      {
        for (int gi = rg.gs; gi < rg.g1; ++gi) {
          const int fb = unit / p.pairs_per_clip;
          const int t0 = 2 * (unit - fb * p.pairs_per_clip);
          mbar_wait(&sempty[slot], ph ^ 1);
          if (p.dbg & 128) {             // timing emulation: no strip loads
            mbar_arrive(&sfull[slot]);
          } else {
            mbar_expect_tx(&sfull[slot], kS2TileStrip);
            tma_load_4d(strips + slot * kS2TileStrip, &mapX, &sfull[slot], 0, 8 * tile, t0 - 2, fb);
          }
          if (++slot == kS2StripSlots) { slot = 0; ph ^= 1; }
          if (++tile == p.tiles) { tile = 0; ++unit; }
        }
      }
    }
  } else if (warp == 8) {
    // =============================================================== MMA issuer
    if (elect_one_sync()) {
      mbar_expect_tx(wbar, 5 * 8192);
      for (int i = 1; i <= 5; ++i) tma_load_2d(stack + i * 8192, &mapW, wbar, (5 - i) * 64, 0);   // block i = W[kt = 5-i]
      mbar_wait(wbar, 0);
      constexpr uint32_t idesc = umma_idesc_bf16(128, kS2N);
      const uint64_t ad0 = umma_desc_sw128_kmajor(smem_u32(stack));
      const uint64_t bd0 = umma_desc_noswizzle_kmajor(smem_u32(ubuf), kS2Plane, 128);
      const uint64_t pd0 = umma_desc_sw128_kmajor(smem_u32(pairb));
      // kh = 6: odd-row plane, unfolded row oy + 3; the second K half is the same place one stage further
      const uint64_t bp0 = umma_desc_noswizzle_kmajor(smem_u32(ubuf) + 3 * kS2RowBytes, kS2UStage, 128);
      uint32_t ph = 0, acc_phase = 0;
      int acc = 0;
      const Stem2Range rg = stem2_range(p);
      {
        for (int gi = rg.gs; gi < rg.g1; ++gi) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 256;
#pragma unroll
          for (int st = 0; st < kS2Stages; ++st) {
            mbar_wait(&ufull[st], ph);
            tc_fence_after();
            // input frame t0 + st - 2: kt = st for frame t0 (rows 0..63), kt = st - 1 for frame t0+1 (rows 64..127)
            // = stack blocks 5-st, 6-st
            const uint64_t adesc = ad0 + (uint64_t)(((5 - st) * 8192) >> 4);
            const uint64_t bdesc = bd0 + (uint64_t)((st * kS2UStage) >> 4);
            if (p.dbg & 1024) {          // all four K steps of every stage (the 8th window-row slot multiplies zero weights)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_bf16(d, adesc + 2 * j, bdesc + (uint64_t)((j * kS2RowBytes) >> 4), idesc, (st | j) ? 1u : 0u);
              umma_commit(&uempty[st]);
            } else {
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                if ((p.dbg & 8) && j) break;
                umma_bf16(d, adesc + 2 * j, bdesc + (uint64_t)((j * kS2RowBytes) >> 4), idesc, (st | j) ? 1u : 0u);
              }
              if (st & 1) {              // window row 6 of stages st-1 and st in one K = 16 step
                umma_bf16(d, pd0 + 2 * (st >> 1), bp0 + (uint64_t)(((st - 1) * kS2UStage) >> 4), idesc, 1u);
                umma_commit(&uempty[st - 1]);
                umma_commit(&uempty[st]);
              }
            }
          }
          ph ^= 1;
          umma_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // =============================================================== epilogue: 8 warps = 4 lane quarters x 2 column halves
    // mode of this warp from its 32 channels (frame g = (warp & 3) >> 1, channels 32 * (warp & 1) + lane)
    const int chn = (warp & 1) * 32 + lane;
    const bool mono = __all_sync(0xffffffffu, __ldg(p.slope + chn) >= 0.f) && !(p.dbg & 2048);   // dbg 2048: general mode
    const bool pos = __all_sync(0xffffffffu, __ldg(p.scale + chn) >= 0.f);
    const int mode = !mono ? 0 : ((pos && !(p.dbg & 4096)) ? 1 : 2);                              // dbg 4096: mode 2 for all
    const int q = warp & 3;
    if (warp >> 2) {
      if (mode == 0) stem2_epilogue<1, 0>(p, tmem_base, tfull, tempty, stg, q, lane);
      else if (mode == 1) stem2_epilogue<1, 1>(p, tmem_base, tfull, tempty, stg, q, lane);
      else stem2_epilogue<1, 2>(p, tmem_base, tfull, tempty, stg, q, lane);
    } else {
      if (mode == 0) stem2_epilogue<0, 0>(p, tmem_base, tfull, tempty, stg, q, lane);
      else if (mode == 1) stem2_epilogue<0, 1>(p, tmem_base, tfull, tempty, stg, q, lane);
      else stem2_epilogue<0, 2>(p, tmem_base, tfull, tempty, stg, q, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<kS2TmemCols>(tmem_base);
  }
}

}  // namespace dl
