// K2: video stem -- Conv3d(1->64, k(5,7,7), s(1,2,2), p(2,3,3)) + BatchNorm3d + PReLU + MaxPool3d((1,3,3),(1,2,2),
// (0,1,1)) fused into one persistent tcgen05 kernel that emits channels-last per-frame maps (so the reference's
// NCTHW -> (N*T)CHW copy, models/video_models/model.py:9-13, disappears).  Optionally reads raw uint8 crops and
// applies the reference preprocessing (x/255, centre crop, (x-mean)/std; dataloaders.py:19-24) in the load.
//
// GEMM view per frame: M = Ho*Wo conv pixels (row-major, 128 per tile), N = 64, K = 5 temporal taps x 64
// (7 rows x 8 columns of the 7x7 window, zero-weight padding) -> one 64-wide K block per temporal tap.
// With a single input channel TMA im2col cannot form operand A (16-byte minimum inner extent), so 8 builder
// warps assemble the A rows in registers from a TMA-staged input strip and write them with tcgen05.st into TENSOR
// MEMORY: the MMAs read A from TMEM (tcgen05.mma [d], [a], b-desc) and only the 2 KB weight slice per MMA from
// shared memory.  With N = 64 an SS MMA is bound by its shared-memory operand fetch (4 KB of A per 32 cycles of
// math, measured ~120 cycles); moving A off the shared-memory port also frees it for the builders' strip reads
// and the pooling epilogue.  A single thread issues the MMAs into a double-buffered TMEM accumulator; 8 epilogue
// warps apply BN+PReLU, park the bf16 conv rows in a shared-memory ring and max-pool completed rows straight to
// global memory.
//
// Roofline: tensor pipe; algorithmic work 2*Ho*Wo*64*245 flop per frame (DESIGN.md "Kernels").
#include "dl_host.cuh"
#include "dl_ptx.cuh"
#include "stem_prepass.cuh"
#include "stem2_conv3d.cuh"

namespace dl {

constexpr int kStemAStages = 6;                     // = stages per (frame pair, tile): stage index == step index, so the MMA thread's
                                                    // descriptors are loop constants; builder group g owns stages g, g+2, g+4; 32 TMEM columns each
constexpr int kStemTmemCols = 512;                  // 2 x 128 accumulator columns (frame pair) + kStemAStages x 32 operand-A columns
constexpr int kStemTmemA = 256;                     // first operand-A column
constexpr int kStemBBytes = 5 * 64 * 64 * 2;        // weights: 5 K blocks of [64 cout x 64 K]
constexpr int kStemThreads = 18 * 32;               // 8 epilogue + 1 MMA + 1 TMA + 8 builder warps
constexpr int kStemEpiThreads = 256;
constexpr int kStripSlotsMax = 16;                  // strip ring (even count, p.strip_slots): slot s % slots -> consumer group s % 2
constexpr int kStemBars = 2 * kStemAStages + 5 + 2 * kStripSlotsMax;
constexpr int kStripRowsMax = 24;                   // strip rows per stage (TMA box height)

struct StemParams {
  int B, T, H, W;
  int Ho, Wo, Hp, Wp, Mf, tiles_per_frame;
  int ring_rows;        // conv rows parked for pooling (any count >= span + 6)
  int strip_rows, strip_pitch;                      // pitch = roundup8(W + 8) elements (cols c = ix + 3)
  const float* scale;
  const float* shift;
  const float* slope;
  uint16_t* y;
  int pairs_per_clip, units;                        // work unit = two consecutive output frames of one clip
  int out_img_rows;     // row pitch of one output frame (>= Hp)
  int strip_slots;
  int dbg;
};

// Two consecutive output frames (t0, t0+1) of one clip are computed together: their 3-D windows share 4 of the 6
// input frames t0-2 .. t0+3, so an A tile (input frame f, spatial tile) is built ONCE and multiplied by the weight
// slices of both frames in one N=128 MMA (D columns [frame t0 | frame t0+1], B rows [W[kt] ; W[kt-1]]).  That halves
// the MMA count and cuts the builders' work to 6/10, and lifts the N=64 MMAs off their operand-A delivery bound
// (~90 cycles for 32 cycles of math).  Stage order: the two stages that touch ONE accumulator come first and
// overwrite it (k = 0), the four shared ones accumulate.
constexpr int kStemSt = 6;                          // input frames (pipeline stages) per (frame pair, tile)
__device__ __forceinline__ int stem_stage_dt(int st) { return st == 0 ? -2 : (st == 1 ? 3 : st - 3); }

// (frame pair, tile, stage) of a pipeline step; every role walks the same sequence.
struct StemCursor {
  int unit, tile, st, stride, tiles;
  __device__ StemCursor(int first, int stride_, int tiles_) : unit(first), tile(0), st(0), stride(stride_), tiles(tiles_) {}
  __device__ __forceinline__ void advance() {
    if (++st == kStemSt) {
      st = 0;
      if (++tile == tiles) { tile = 0; unit += stride; }
    }
  }
};


__global__ void __launch_bounds__(kStemThreads, 1)
stem_conv3d_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
                   const StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // round up inside the shared window (pointer arithmetic on the __shared__ symbol keeps LDS/STS addressing)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smB = smem;                                           // 40 KB: block i = W[kt = 4 - i]
  uint8_t* ring = smB + kStemBBytes;                             // 2 frames x ring_rows x Wo x 128 B
  const int ring_bytes = p.ring_rows * p.Wo * 128;
  uint16_t* strip = reinterpret_cast<uint16_t*>(ring + 2 * ring_bytes);   // strip_slots x strip_rows x strip_pitch bf16
  const int strip_elems = p.strip_rows * p.strip_pitch;
  const int strip_buf = (strip_elems + 63) & ~63;
  float* chan = reinterpret_cast<float*>(strip + p.strip_slots * strip_buf);   // scale, shift, slope
  uint64_t* bars = reinterpret_cast<uint64_t*>(chan + 192);
  uint64_t* full = bars;                         // [kStemAStages]
  uint64_t* empty = bars + kStemAStages;         // [kStemAStages]
  uint64_t* tfull = bars + 2 * kStemAStages;     // [2]
  uint64_t* tempty = tfull + 2;                  // [2]
  uint64_t* wbar = tempty + 2;                   // [1]
  uint64_t* sfull = wbar + 1;                    // [kStripSlotsMax] strip slot filled (TMA)
  uint64_t* sempty = sfull + kStripSlotsMax;     // [kStripSlotsMax] strip slot consumed (4 builder warps of one group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + kStripSlotsMax);
  int* tile_iy0 = reinterpret_cast<int*>(sempty + kStripSlotsMax + 1);   // [tiles_per_frame <= 64] first input row of a tile's strip

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 192; i += kStemThreads) {
    const int c = i & 63;
    chan[i] = i < 64 ? p.scale[c] : (i < 128 ? p.shift[c] : p.slope[c]);
  }
  for (int i = threadIdx.x; i < p.tiles_per_frame; i += kStemThreads) tile_iy0[i] = 2 * ((i * 128) / p.Wo) - 3;
  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < kStemAStages; ++s) {
        mbar_init(&full[s], 4);      // one elected arrive per builder warp of the owning group
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < p.strip_slots; ++s) {
        mbar_init(&sfull[s], 1);      // the producer's expect_tx arrive
        mbar_init(&sempty[s], 4);
      }
      mbar_init(&tfull[0], 1);
      mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], kStemEpiThreads);
      mbar_init(&tempty[1], kStemEpiThreads);
      mbar_init(wbar, 1);
      fence_mbar_init();
      tma_prefetch_desc(&mapW);
      tma_prefetch_desc(&mapX);
    }
    __syncwarp();
    tmem_alloc<kStemTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 10) {
    // =============================================================== builders: operand A from the staged strip
    // Two groups of 4 warps alternate pipeline stages (group g owns stages s = g, g+2, ...), so two A tiles are
    // in flight; thread <-> A row (conv pixel) = TMEM lane (a warp may only touch lane quarter warp & 3); the row is
    // 8 chunks of 16 B: chunk kh = 8 consecutive input pixels of window row kh (chunk 7 = zero padding of K).
    const int group = (warp - 10) >> 2;
    const int arow = (warp & 3) * 32 + lane;
    const int SP = p.strip_pitch;
    StemCursor cur(blockIdx.x, gridDim.x, p.tiles_per_frame);
    if (group == 1) cur.advance();
    int cached_tile = -1;
    uint32_t arel = 0;
    bool avalid = false;
    uint32_t sslot = group, sph = 0;              // strip ring position of step s: s % strip_slots, (s / strip_slots) & 1
    int slot = group;                             // A stage of step s: s % kStemAStages, phase (s / kStemAStages) & 1
    uint32_t ph = 0;
    for (; cur.unit < p.units;) {
      if (cur.tile != cached_tile) {
        cached_tile = cur.tile;
        const int m = cur.tile * 128 + arow;
        avalid = m < p.Mf;
        const int yy = m / p.Wo, xx = m - yy * p.Wo;
        arel = (uint32_t)((2 * yy - 3 - tile_iy0[cur.tile]) * SP + 2 * xx) >> 1;   // uint32 index into the strip
      }
      const uint32_t* srow = reinterpret_cast<const uint32_t*>(strip + sslot * strip_buf) + arel;
      mbar_wait(&sfull[sslot], sph);
      uint4 v[8];
#pragma unroll
      for (int kh = 0; kh < 7; ++kh) {
        v[kh] = make_uint4(0u, 0u, 0u, 0u);
        if (avalid && !(p.dbg & 16)) {
          const uint32_t* sp = srow + kh * (SP >> 1);
          v[kh].x = sp[0]; v[kh].y = sp[1]; v[kh].z = sp[2]; v[kh].w = sp[3];
        }
      }
      v[7] = make_uint4(0u, 0u, 0u, 0u);
      mbar_wait(&empty[slot], ph ^ 1);              // the MMAs that read this A stage have completed
      tc_fence_after();
      if (!(p.dbg & 16)) {
        uint32_t a[32];                             // K order kh*8 + kw = the order of the packed weights
#pragma unroll
        for (int kh = 0; kh < 8; ++kh) { a[4 * kh] = v[kh].x; a[4 * kh + 1] = v[kh].y; a[4 * kh + 2] = v[kh].z; a[4 * kh + 3] = v[kh].w; }
        tmem_st_32x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kStemTmemA + slot * 32, a);
        tmem_st_wait();
      }
      tc_fence_before();
      // The strip slot is released to TMA below.  Its reads have been PERFORMED by now -- tcgen05.st consumed the
      // loaded registers -- so no generic->async proxy fence is needed (an arrive issued right after bare loads
      // could overtake them).
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sempty[sslot]);
        mbar_arrive(&full[slot]);
      }
      sslot += 2;
      if (sslot >= (uint32_t)p.strip_slots) { sslot -= p.strip_slots; sph ^= 1; }
      slot += 2;
      if (slot >= kStemAStages) { slot -= kStemAStages; ph ^= 1; }
      cur.advance();
      cur.advance();
    }
  } else if (warp == 9) {
    // =============================================================== strip producer: one thread, TMA only
    // The input was normalised / zero-bordered once by stem_prepass_kernel into (B, T, H+8, pitch) bf16, so the
    // strip of a stage is a plain 2-D box of that tensor: rows iy0 .. iy0+strip_rows-1 of input frame t0+dt (a frame
    // index outside [0,T) is out of bounds in the T dimension -> the TMA unit zero-fills it: Conv3d's temporal pad).
    if (elect_one_sync()) {
      StemCursor cur(blockIdx.x, gridDim.x, p.tiles_per_frame);
      int last_unit = -1, fb = 0, t0 = 0;
      const uint32_t strip_bytes = (uint32_t)strip_elems * 2u;
      uint32_t slot = 0, ph = 0;
      for (; cur.unit < p.units; cur.advance()) {
        if (cur.unit != last_unit) { last_unit = cur.unit; fb = cur.unit / p.pairs_per_clip; t0 = 2 * (cur.unit - fb * p.pairs_per_clip); }
        mbar_wait(&sempty[slot], ph ^ 1);
        if (p.dbg & 16384) {             // timing emulation: no strip load at all (is the TMA unit's row rate the pace?)
          mbar_arrive(&sfull[slot]);
        } else {
          mbar_expect_tx(&sfull[slot], strip_bytes);
          tma_load_4d(strip + slot * strip_buf, &mapX, &sfull[slot], 0, tile_iy0[cur.tile] + 3, t0 + stem_stage_dt(cur.st), fb);
        }
        if (++slot == (uint32_t)p.strip_slots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 8) {
    // =============================================================== MMA issuer
    if (elect_one_sync()) {
      mbar_expect_tx(wbar, kStemBBytes);
      for (int i = 0; i < 5; ++i) tma_load_2d(smB + i * 8192, &mapW, wbar, (4 - i) * 64, 0);   // block i = W[kt = 4-i]
      mbar_wait(wbar, 0);
      constexpr uint32_t idesc64 = umma_idesc_bf16(128, 64);
      constexpr uint32_t idesc128 = umma_idesc_bf16(128, 128);
      static_assert(kStemAStages == kStemSt, "stage index == step index");
      // A single thread issues everything, so every instruction between two MMAs is pipe idle time: all operand
      // addresses below are loop invariants (the compiler keeps them in uniform registers).
      const uint64_t bd0 = umma_desc_sw128_kmajor(smem_u32(smB));
      const uint32_t a0 = tmem_base + kStemTmemA;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        for (int tile = 0; tile < p.tiles_per_frame; ++tile) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 128;           // columns [frame t0 | frame t0+1]
#pragma unroll
          for (int st = 0; st < kStemSt; ++st) {
            mbar_wait(&full[st], phase);
            tc_fence_after();
            // st 0: input t0-2 -> frame t0 only (kt 0);  st 1: input t0+3 -> frame t0+1 only (kt 4);
            // st 2..5: input t0+dt, kt = dt+2 for frame t0 and kt-1 for frame t0+1 = adjacent weight blocks
            const uint32_t dd = st == 1 ? d + 64 : d;
            const uint32_t wblk = st == 0 ? 4u : (st == 1 ? 0u : (uint32_t)(5 - st));
            const uint64_t bdesc = bd0 + (uint64_t)(wblk * 8192u >> 4);
            // dbg 4096 / 8192 (timing emulation only, results are garbage): issue the shared stages at N = 192 / 256 to
            // measure what an MMA with A in tensor memory costs as N grows
            const uint32_t idesc = st < 2 ? idesc64 : ((p.dbg & 4096) ? umma_idesc_bf16(128, 192) : ((p.dbg & 8192) ? umma_idesc_bf16(128, 256) : idesc128));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if ((p.dbg & 8) && k) break;
              umma_bf16_ts(dd, a0 + st * 32 + 8 * k, bdesc + 2 * k, idesc, (st >= 2 || k != 0) ? 1u : 0u);   // 8 columns per K = 16
            }
            umma_commit(&empty[st]);
          }
          phase ^= 1;
          umma_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // =============================================================== epilogue: BN + PReLU -> ring -> max-pool
    // 8 warps: warp & 3 = TMEM lane quarter (tile rows 32*(warp&3) ..), warp >> 2 = channel half (32 channels);
    // both frames of the pair per tile, each with its own ring of conv rows.
    const int et = threadIdx.x;                 // 0..255
    const int erow = et & 127;
    const int half = et >> 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int row_bytes = p.Wo * 128;
    // this thread's conv pixel (yy, xx) and ring row advance by 128 pixels per tile: no divisions in the loop
    const int yy0 = erow / p.Wo, xx0 = erow - yy0 * p.Wo, ry0 = yy0 % p.ring_rows;
    const int sy = 128 / p.Wo, sx = 128 - sy * p.Wo;
    const int per_row = p.Wp * 8;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int fb = unit / p.pairs_per_clip;
      const int t0 = 2 * (unit - fb * p.pairs_per_clip);
      const int nfr = min(2, p.T - t0);                     // a clip with an odd frame count ends on a half pair
      int py_done = 0, pr1 = 0;                             // pr1 = (2 * py_done) % ring_rows
      int yy = yy0, xx = xx0, ry = ry0;
      uint16_t* yframe0 = p.y + ((size_t)fb * p.T + t0) * p.out_img_rows * p.Wp * 64;
      for (int tile = 0; tile < p.tiles_per_frame; ++tile) {
        const int m = tile * 128 + erow;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        uint32_t r[2][32];
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * 128 + half * 32;
        tmem_ld_32x32(taddr, r[0]);
        tmem_ld_32x32(taddr + 64, r[1]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        if (m < p.Mf && !(p.dbg & (32 | 2048))) {
          const float4* c4 = reinterpret_cast<const float4*>(chan) + half * 8;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint8_t* dst = ring + g * ring_bytes + ry * row_bytes + xx * 128;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              float v[8];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 sc = c4[2 * ch + h], sh = c4[16 + 2 * ch + h], sl = c4[32 + 2 * ch + h];
                const int c = ch * 8 + 4 * h;
                float z;
                z = fmaf(__uint_as_float(r[g][c]), sc.x, sh.x);         v[4 * h + 0] = z > 0.f ? z : z * sl.x;
                z = fmaf(__uint_as_float(r[g][c + 1]), sc.y, sh.y);     v[4 * h + 1] = z > 0.f ? z : z * sl.y;
                z = fmaf(__uint_as_float(r[g][c + 2]), sc.z, sh.z);     v[4 * h + 2] = z > 0.f ? z : z * sl.z;
                z = fmaf(__uint_as_float(r[g][c + 3]), sc.w, sh.w);     v[4 * h + 3] = z > 0.f ? z : z * sl.w;
              }
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(dst + (((half * 4 + ch) ^ (xx & 7)) << 4)) = o;
            }
          }
        }
        named_bar_sync(2, kStemEpiThreads);
        // pooled rows whose three conv rows are now complete
        const int m_end = min((tile + 1) * 128, p.Mf);
        const int rows_complete = m_end / p.Wo;               // conv rows 0 .. rows_complete-1 are final
        const int py_ready = rows_complete / 2;               // needs conv row 2*py+1 <= rows_complete-1
        xx += sx; yy += sy; ry += sy;
        if (xx >= p.Wo) { xx -= p.Wo; ++yy; ++ry; }
        if (ry >= p.ring_rows) ry -= p.ring_rows;
        // work items (frame g, pooled row py, pixel px, 16-byte channel chunk ch) of this tile spread evenly over the
        // 256 threads; decoded by subtraction (a tile completes at most a few pooled rows)
        const int npy = py_ready - py_done;
        const int n1 = npy * per_row;
        for (int it = et; it < n1 * nfr && !(p.dbg & (32 | 1024)); it += kStemEpiThreads) {
          const int g = it >= n1 ? 1 : 0;
          int rem = it - g * n1;
          int py = py_done, q1 = pr1;                          // q1 = (2 * py) % ring_rows
          while (rem >= per_row) {
            rem -= per_row; ++py; q1 += 2;
            if (q1 >= p.ring_rows) q1 -= p.ring_rows;
          }
          // ring rows of conv rows 2py-1 (clamped: repeating a row leaves the max unchanged), 2py, 2py+1
          const int q0 = py > 0 ? (q1 == 0 ? p.ring_rows - 1 : q1 - 1) : 0;
          const int q2 = q1 + 1 == p.ring_rows ? 0 : q1 + 1;
          const int ch = rem & 7, px = rem >> 3;
          {
          const uint8_t* rg = ring + g * ring_bytes;
          const uint8_t* r0p = rg + q0 * row_bytes;
          const uint8_t* r1p = rg + q1 * row_bytes;
          const uint8_t* r2p = rg + q2 * row_bytes;
          const int cx1 = 2 * px, cx2 = cx1 + 1, cx0 = max(cx1 - 1, 0);
          const int o0 = cx0 * 128 + ((ch ^ (cx0 & 7)) << 4);
          const int o1 = cx1 * 128 + ((ch ^ (cx1 & 7)) << 4);
          const int o2 = cx2 * 128 + ((ch ^ (cx2 & 7)) << 4);
          uint4 v[9];
          v[0] = *reinterpret_cast<const uint4*>(r0p + o0); v[1] = *reinterpret_cast<const uint4*>(r0p + o1);
          v[2] = *reinterpret_cast<const uint4*>(r0p + o2); v[3] = *reinterpret_cast<const uint4*>(r1p + o0);
          v[4] = *reinterpret_cast<const uint4*>(r1p + o1); v[5] = *reinterpret_cast<const uint4*>(r1p + o2);
          v[6] = *reinterpret_cast<const uint4*>(r2p + o0); v[7] = *reinterpret_cast<const uint4*>(r2p + o1);
          v[8] = *reinterpret_cast<const uint4*>(r2p + o2);
          __nv_bfloat162 best[4];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            __nv_bfloat162 m01 = __hmax2(reinterpret_cast<const __nv_bfloat162*>(&v[0])[qq],
                                         reinterpret_cast<const __nv_bfloat162*>(&v[1])[qq]);
            __nv_bfloat162 m23 = __hmax2(reinterpret_cast<const __nv_bfloat162*>(&v[2])[qq],
                                         reinterpret_cast<const __nv_bfloat162*>(&v[3])[qq]);
            __nv_bfloat162 m45 = __hmax2(reinterpret_cast<const __nv_bfloat162*>(&v[4])[qq],
                                         reinterpret_cast<const __nv_bfloat162*>(&v[5])[qq]);
            __nv_bfloat162 m67 = __hmax2(reinterpret_cast<const __nv_bfloat162*>(&v[6])[qq],
                                         reinterpret_cast<const __nv_bfloat162*>(&v[7])[qq]);
            best[qq] = __hmax2(__hmax2(__hmax2(m01, m23), __hmax2(m45, m67)),
                               reinterpret_cast<const __nv_bfloat162*>(&v[8])[qq]);
          }
          uint16_t* yf = yframe0 + (size_t)g * p.out_img_rows * p.Wp * 64;
          *reinterpret_cast<uint4*>(yf + ((size_t)py * p.Wp + px) * 64 + ch * 8) = *reinterpret_cast<const uint4*>(best);
          }
        }
        pr1 = (pr1 + 2 * npy) % p.ring_rows;
        py_done = py_ready;
        // No second barrier: a warp that runs ahead writes the NEXT tile's conv rows [rc, rc+span] while the others
        // still pool rows [rc-5, rc-1] -- disjoint in a ring of span + 6 rows -- and then stops at that tile's barrier.
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<kStemTmemCols>(tmem_base);
  }
}

// Pre-pass (HBM-bound, ~0.03 ms per 64 utterances): uint8 crops (x/255, centre crop, (x-mean)/std fused;
// models/video_models/dataloaders.py:19-24) or normalised f32 frames -> zero-bordered bf16 frames
// xp (B*T, H+8, pitch): row iy+3, column ix+3.  One thread writes 8 consecutive columns (16 B).
__global__ void stem_prepass_kernel(const void* __restrict__ x, int is_u8, int frames, int H, int W, int Hraw,
                                    int Wraw, int dh, int dw, float u8_scale, float u8_bias, int rows, int pitch,
                                    uint16_t* __restrict__ xp) {
  const int groups = pitch >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)frames * rows * groups) return;
  const int g = (int)(idx % groups);
  const long long t = idx / groups;
  const int row = (int)(t % rows);
  const long long f = t / rows;
  const int iy = row - 3;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ix = g * 8 + j - 3;
    float val = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      if (is_u8) {
        const uint8_t u = __ldg(static_cast<const uint8_t*>(x) + ((size_t)f * Hraw + (iy + dh)) * Wraw + (ix + dw));
        val = fmaf((float)u, u8_scale, u8_bias);
      } else {
        val = __ldg(static_cast<const float*>(x) + ((size_t)f * H + iy) * W + ix);
      }
    }
    v[j] = val;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(xp + ((size_t)f * rows + row) * pitch + g * 8) = o;
}

}  // namespace dl

extern "C" long long dl_stem_workspace_bytes(int B, int T, int H, int W) {
  if (B <= 0 || T <= 0 || H <= 0 || W <= 0) return 0;
  const long long pitch = (W + 8 + 7) / 8 * 8;
  return (long long)B * T * (H + 8) * pitch * 2;
}

// phases: 1 = pre-pass only (x -> workspace), 2 = main kernel only (workspace -> y), 3 = both
static int stem_impl(int phases, const void* x, int is_u8, int B, int T, int H, int W, int Hraw, int Wraw,
                     float mean, float std, const void* w_packed, const float* scale,
                     const float* shift, const float* slope, void* y, int out_img_rows,
                     const int32_t* lengths, void* workspace, void* stream) {
  using namespace dl;
  DL_CHECK_ARG(((phases & 1) == 0 || x) && ((phases & 2) == 0 || (w_packed && scale && shift && slope && y)) && workspace,
               "stem: null pointer");
  DL_CHECK_ARG(B > 0 && T > 0, "stem: empty batch");
  DL_CHECK_ARG(H >= 8 && W >= 32 && H % 4 == 0 && W % 4 == 0 && W <= 120, "stem: H, W must be multiples of 4, 32 <= W <= 120");
  if (is_u8) {
    DL_CHECK_ARG(Hraw >= H && Wraw >= W && std != 0.f, "stem: raw crop smaller than the centre crop");
  }
  int st = require_sm100();
  if (st != DL_OK) return st;

  StemParams p;
  p.B = B; p.T = T; p.H = H; p.W = W;
  p.Ho = H / 2; p.Wo = W / 2; p.Hp = H / 4; p.Wp = W / 4;
  p.Mf = p.Ho * p.Wo;
  p.tiles_per_frame = (p.Mf + 127) / 128;
  const int span = (127 + p.Wo - 1) / p.Wo;           // extra conv rows a tile can reach past its first row
  p.ring_rows = span + 6;
  p.strip_rows = 2 * span + 7;
  p.strip_pitch = (W + 8 + 7) / 8 * 8;
  DL_CHECK_ARG(p.strip_rows <= kStripRowsMax, "stem: frame too narrow (strip of %d rows)", p.strip_rows);
  p.scale = scale; p.shift = shift; p.slope = slope;
  p.y = static_cast<uint16_t*>(y);
  p.pairs_per_clip = (T + 1) / 2;
  p.units = B * p.pairs_per_clip;
  p.out_img_rows = out_img_rows > 0 ? out_img_rows : p.Hp;
  p.dbg = opt_dbg();
  p.strip_slots = 6;
  {   // measurement aid: dbg bits 16..19 = strip ring slots / 2 (even counts 2..16)
    const int s2 = (p.dbg >> 16) & 15;
    if (s2 >= 1 && 2 * s2 <= kStripSlotsMax) p.strip_slots = 2 * s2;
  }
  DL_CHECK_ARG(p.out_img_rows >= p.Hp, "stem: out_img_rows < H/4");
  DL_CHECK_ARG(p.tiles_per_frame <= 64, "stem: frame too large (more than 64 tiles)");

  cudaStream_t cs = (cudaStream_t)stream;
  // ---- pre-pass: normalise + zero-border into the caller's workspace
  const int rows = H + 8, pitch = p.strip_pitch;
  // CenterCrop: delta = int(round(w - tw) / 2.)  (models/video_models/preprocess.py:88-90)
  const int dh = is_u8 ? (Hraw - H) / 2 : 0, dw = is_u8 ? (Wraw - W) / 2 : 0;
  if (!(phases & 1)) {
    // the caller ran dl_stem_prepass on this workspace already (e.g. on another stream, under the audio branch)
  } else if (opt_prepass() >= 2 || lengths) {
    const int aligned4 = (Wraw % 4 == 0 && ((uintptr_t)x & 3) == 0) ? 1 : 0;
    stem_prepass2_kernel<<<(unsigned)(B * T), 256, 0, cs>>>(
        x, is_u8, H, W, Hraw, Wraw, dh, dw, is_u8 ? 1.0f / (255.0f * std) : 1.0f, is_u8 ? -mean / std : 0.0f, rows,
        pitch, aligned4, T, lengths, static_cast<uint16_t*>(workspace));
    st = check_launch("stem_prepass2_kernel");
    if (st != DL_OK) return st;
  } else {
    const long long n = (long long)B * T * rows * (pitch / 8);
    stem_prepass_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(
        x, is_u8, B * T, H, W, Hraw, Wraw, dh, dw, is_u8 ? 1.0f / (255.0f * std) : 1.0f, is_u8 ? -mean / std : 0.0f,
        rows, pitch, static_cast<uint16_t*>(workspace));
    st = check_launch("stem_prepass_kernel");
    if (st != DL_OK) return st;
  }
  if (!(phases & 2)) return DL_OK;

  if (opt_stem() >= 2 && W == 2 * kS2Wo && H % 8 == 0) {
    // ---- second generation (stem2_conv3d.cuh): channels on lanes, four conv rows per tile
    Stem2Params q;
    q.B = B; q.T = T;
    q.tiles = H / 8;
    q.pairs_per_clip = (T + 1) / 2;
    q.units = B * q.pairs_per_clip;
    q.out_img_rows = p.out_img_rows;
    q.scale = scale; q.shift = shift; q.slope = slope;
    q.y = static_cast<uint16_t*>(y);
    q.w = static_cast<const uint16_t*>(w_packed);
    q.dbg = p.dbg;
    CUtensorMap mapW2, mapX2;
    st = make_tiled_2d_bf16(&mapW2, w_packed, 64, 320, 320, 64, 64);
    if (st != DL_OK) return st;
    st = make_tiled_4d_bf16_noswizzle(&mapX2, workspace, (uint64_t)pitch, (uint64_t)rows, (uint64_t)T, (uint64_t)B,
                                      (uint32_t)pitch, (uint32_t)kS2StripRows, 6);
    if (st != DL_OK) return st;
    int grid2 = device_sm_count();
    if (grid2 <= 0) grid2 = 148;
    if (q.units * q.tiles < grid2) grid2 = q.units * q.tiles;
    static PerDevice<int> configured2_dev;
    int* configured2 = configured2_dev.slot();
    if (!configured2) return fail(DL_ERR_CUDA, "stem: no current device");
    if (!*configured2) {
      cudaError_t e = cudaFuncSetAttribute(stem2_conv3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kS2Smem);
      if (e != cudaSuccess) return fail(DL_ERR_CUDA, "stem2 smem attribute: %s", cudaGetErrorString(e));
      *configured2 = 1;
    }
    stem2_conv3d_kernel<<<grid2, kS2Threads, kS2Smem, cs>>>(mapW2, mapX2, q);
    return check_launch("stem2_conv3d_kernel");
  }

  const int strip_elems = p.strip_rows * p.strip_pitch;
  const size_t smem = 1024 + (size_t)kStemBBytes + 2 * (size_t)p.ring_rows * p.Wo * 128 +
                      p.strip_slots * (size_t)((strip_elems + 63) & ~63) * 2 + 192 * 4 + (kStemBars + 1) * 8 + 16 + 64 * 4;
  DL_CHECK_ARG(smem <= 227 * 1024, "stem: shared-memory budget exceeded (%zu B)", smem);
  CUtensorMap mapW, mapX;
  st = make_tiled_2d_bf16(&mapW, w_packed, 64, 320, 320, 64, 64);
  if (st != DL_OK) return st;
  st = make_tiled_4d_bf16_noswizzle(&mapX, workspace, (uint64_t)pitch, (uint64_t)rows, (uint64_t)T, (uint64_t)B,
                                    (uint32_t)pitch, (uint32_t)p.strip_rows);
  if (st != DL_OK) return st;
  int grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (p.units < grid) grid = p.units;
  static PerDevice<size_t> configured_dev;
  size_t* configured = configured_dev.slot();
  if (!configured) return fail(DL_ERR_CUDA, "stem: no current device");
  if (smem > *configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_conv3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "stem smem attribute: %s", cudaGetErrorString(e));
    *configured = smem;
  }
  stem_conv3d_kernel<<<grid, kStemThreads, smem, cs>>>(mapW, mapX, p);
  return check_launch("stem_conv3d_kernel");
}

extern "C" int dl_stem_conv3d_bn_prelu_pool(const void* x, int is_u8, int B, int T, int H, int W, int Hraw, int Wraw,
                                            float mean, float std, const void* w_packed, const float* scale,
                                            const float* shift, const float* slope, void* y, int out_img_rows,
                                            const int32_t* lengths, void* workspace, void* stream) {
  return stem_impl(3, x, is_u8, B, T, H, W, Hraw, Wraw, mean, std, w_packed, scale, shift, slope, y, out_img_rows, lengths,
                   workspace, stream);
}

extern "C" int dl_stem_prepass(const void* x, int is_u8, int B, int T, int H, int W, int Hraw, int Wraw, float mean,
                               float std, const int32_t* lengths, void* workspace, void* stream) {
  return stem_impl(1, x, is_u8, B, T, H, W, Hraw, Wraw, mean, std, nullptr, nullptr, nullptr, nullptr, nullptr, 0, lengths,
                   workspace, stream);
}

extern "C" int dl_stem_conv3d_prepassed(int B, int T, int H, int W, const void* w_packed, const float* scale,
                                        const float* shift, const float* slope, void* y, int out_img_rows,
                                        void* workspace, void* stream) {
  return stem_impl(2, nullptr, 0, B, T, H, W, H, W, 0.f, 1.f, w_packed, scale, shift, slope, y, out_img_rows, nullptr,
                   workspace, stream);
}
