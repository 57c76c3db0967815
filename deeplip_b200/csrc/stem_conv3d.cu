// K2: video stem -- Conv3d(1->64, k(5,7,7), s(1,2,2), p(2,3,3)) + BatchNorm3d + PReLU + MaxPool3d((1,3,3),(1,2,2),
// (0,1,1)) fused into one persistent tcgen05 kernel that emits channels-last per-frame maps (so the reference's
// NCTHW -> (N*T)CHW copy, models/video_models/model.py:9-13, disappears).  Optionally reads raw uint8 crops and
// applies the reference preprocessing (x/255, centre crop, (x-mean)/std; dataloaders.py:19-24) in the load.
//
// GEMM view per frame: M = Ho*Wo conv pixels (row-major, 128 per tile), N = 64, K = 5 temporal taps x 64
// (7 rows x 8 columns of the 7x7 window, zero-weight padding) -> one 64-wide K block per temporal tap.
// With a single input channel TMA im2col cannot form operand A (16-byte minimum inner extent), so 8 producer
// warps build the 128B-swizzled K-major A tile in shared memory from a small staged input strip; a single
// thread issues tcgen05.mma into a double-buffered TMEM accumulator; 4 epilogue warps apply BN+PReLU, park the
// bf16 conv rows in a shared-memory ring and max-pool completed rows straight to global memory.
//
// Roofline: tensor pipe; algorithmic work 2*Ho*Wo*64*245 flop per frame (DESIGN.md "Kernels").
#include <type_traits>
#include "dl_host.cuh"
#include "dl_ptx.cuh"

namespace dl {

constexpr int kStemAStages = 4;
constexpr int kStemABytes = 128 * 64 * 2;           // one A tile: 128 pixels x 64 K (bf16)
constexpr int kStemBBytes = 5 * 64 * 64 * 2;        // weights: 5 K blocks of [64 cout x 64 K]
constexpr int kStemThreads = 16 * 32;               // 4 epilogue + 1 MMA + 3 loader + 8 builder warps
constexpr int kLoaders = 3;                         // loader warps
constexpr int kStripSlots = 6;                      // strip ring: slot s % 6 -> one producer warp (s % 3), one consumer group (s % 2)
constexpr int kStripRowsMax = 24;                   // loader warp lw fills strip rows lw, lw+2, ...

struct StemParams {
  const void* x;
  int B, T, H, W, Hraw, Wraw, dh, dw;
  float u8_scale, u8_bias;                          // (u/255 - mean)/std == u * u8_scale + u8_bias
  int Ho, Wo, Hp, Wp, Mf, tiles_per_frame;
  int ring_rows;        // power of two
  int strip_rows, strip_pitch;                      // pitch = W + 12 elements (cols c = ix + 3, c in [0, W+8))
  const float* scale;
  const float* shift;
  const float* slope;
  uint16_t* y;
  int frames;
  int out_img_rows;     // row pitch of one output frame (>= Hp)
};

// (frame, tile, temporal tap) of a pipeline stage; every role walks the same sequence.
struct StemCursor {
  int frame, tile, kt, ft, stride, tiles;
  __device__ StemCursor(int first, int stride_, int tiles_) : frame(first), tile(0), kt(0), ft(0), stride(stride_), tiles(tiles_) {}
  __device__ __forceinline__ void advance() {
    if (++kt == 5) {
      kt = 0;
      if (++tile == tiles) { tile = 0; frame += stride; }
    }
  }
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <bool kU8, int kIters>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_conv3d_kernel(const __grid_constant__ CUtensorMap mapW, const StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // round up inside the shared window (pointer arithmetic on the __shared__ symbol keeps LDS/STS addressing)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smA = smem;                                           // kStemAStages x 16 KB
  uint8_t* smB = smA + kStemAStages * kStemABytes;               // 40 KB
  uint8_t* ring = smB + kStemBBytes;                             // ring_rows x Wo x 128 B
  const int ring_bytes = p.ring_rows * p.Wo * 128;
  uint16_t* strip = reinterpret_cast<uint16_t*>(ring + ring_bytes);   // kStripSlots x strip_rows x strip_pitch bf16
  const int strip_elems = p.strip_rows * p.strip_pitch;
  const int strip_buf = (strip_elems + 7) & ~7;
  float* chan = reinterpret_cast<float*>(strip + kStripSlots * strip_buf);   // scale, shift, slope
  uint64_t* bars = reinterpret_cast<uint64_t*>(chan + 192);
  uint64_t* full = bars;                         // [kStemAStages]
  uint64_t* empty = bars + kStemAStages;         // [kStemAStages]
  uint64_t* tfull = bars + 2 * kStemAStages;     // [2]
  uint64_t* tempty = tfull + 2;                  // [2]
  uint64_t* wbar = tempty + 2;                   // [1]
  uint64_t* sfull = wbar + 1;                    // [kStripSlots] strip slot filled (its loader warp)
  uint64_t* sempty = sfull + kStripSlots;        // [kStripSlots] strip slot consumed (4 builder warps of one group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + kStripSlots);
  int* tile_iy0 = reinterpret_cast<int*>(sempty + kStripSlots + 1);   // [tiles_per_frame <= 64] first input row of a tile's strip

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 192; i += kStemThreads) {
    const int c = i & 63;
    chan[i] = i < 64 ? p.scale[c] : (i < 128 ? p.shift[c] : p.slope[c]);
  }
  for (int i = threadIdx.x; i < p.tiles_per_frame; i += kStemThreads) tile_iy0[i] = 2 * ((i * 128) / p.Wo) - 3;
  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < kStemAStages; ++s) {
        mbar_init(&full[s], 4);      // one elected arrive per builder warp of the owning group
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < kStripSlots; ++s) {
        mbar_init(&sfull[s], 1);
        mbar_init(&sempty[s], 4);
      }
      mbar_init(&tfull[0], 1);
      mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 128);
      mbar_init(&tempty[1], 128);
      mbar_init(wbar, 1);
      fence_mbar_init();
      tma_prefetch_desc(&mapW);
    }
    __syncwarp();
    tmem_alloc<128>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    // =============================================================== builders: operand A from the staged strip
    // Two groups of 4 warps alternate pipeline stages (group g owns stages s = g, g+2, ...), so two A tiles are
    // in flight; thread <-> A-tile row (conv pixel), 8 chunks of 16 B: chunk kh = 8 consecutive input pixels of
    // window row kh (chunk 7 = zero padding of K).
    const int bt = threadIdx.x - 8 * 32;          // 0..255
    const int group = bt >> 7;
    const int arow = bt & 127;
    const int SP = p.strip_pitch;
    StemCursor cur(blockIdx.x, gridDim.x, p.tiles_per_frame);
    if (group == 1) cur.advance();
    int cached_tile = -1, cached_frame = -1;
    uint32_t arel = 0;
    bool avalid = false;
    uint32_t sslot = group, sph = 0;              // strip ring position of stage s: s % kStripSlots, (s / kStripSlots) & 1
    for (uint32_t s = group; cur.frame < p.frames; s += 2) {
      const int slot = s & 3;
      const uint32_t ph = (s >> 2) & 1;
      if (cur.tile != cached_tile || cur.frame != cached_frame) {
        cached_tile = cur.tile; cached_frame = cur.frame;
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
        if (avalid) {
          const uint32_t* sp = srow + kh * (SP >> 1);
          v[kh].x = sp[0]; v[kh].y = sp[1]; v[kh].z = sp[2]; v[kh].w = sp[3];
        }
      }
      v[7] = make_uint4(0u, 0u, 0u, 0u);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sempty[sslot]);         // strip slot consumed (values are in registers)
      sslot += 2;
      if (sslot >= kStripSlots) { sslot -= kStripSlots; sph ^= 1; }
      mbar_wait(&empty[slot], ph ^ 1);
      uint8_t* dst_row = smA + slot * kStemABytes + arow * 128;
#pragma unroll
      for (int kh = 0; kh < 8; ++kh) *reinterpret_cast<uint4*>(dst_row + ((kh ^ (arow & 7)) << 4)) = v[kh];
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[slot]);
      cur.advance();
      cur.advance();
    }
  } else if (warp >= 5) {
    // =============================================================== loaders: global -> bf16 strip ring
    // Three loader warps; warp lw owns pipeline stages s = lw, lw+3, ... (strip slots lw and lw+3).  Lane <-> group of 4
    // input columns ix0 = 4*lane-4 .. +3 (one aligned 4-byte / 16-byte global load per strip row).  The loads of a
    // warp's next stage are issued before it converts the current one, so ~6 stages of global loads are in
    // flight per CTA and their latency never reaches the builders.
    const int lw = warp - 5;
    const int grp = lane;
    const int ix0 = 4 * grp - 4;
    const bool col_ok = ix0 >= 0 && ix0 < p.W;            // whole group inside the image (W % 4 == 0)
    const bool grp_ok = grp <= p.W / 4 + 2;               // group touches the strip at all
    const int SP = p.strip_pitch;
    const size_t frame_elems = (size_t)p.Hraw * p.Wraw;
    const int col_off = (kU8 ? p.dh * p.Wraw + p.dw : 0) + ix0;
    typename std::conditional<kU8, uint32_t, float4>::type raw[kIters];
    uint32_t vmask = 0;

    auto issue = [&](const StemCursor& c) {
      const int tt = c.ft + c.kt - 2;
      vmask = 0;
      if (!(col_ok && tt >= 0 && tt < p.T)) return;
      const int iy0 = tile_iy0[c.tile];
      const size_t fo = (size_t)(c.frame + c.kt - 2) * frame_elems + col_off;
#pragma unroll
      for (int r = 0; r < kIters; ++r) {
        const int iy = iy0 + r;
        if (r < p.strip_rows && (unsigned)iy < (unsigned)p.H) {
          if constexpr (kU8) {
            raw[r] = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.x) + fo + iy * p.Wraw));
          } else {
            raw[r] = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(p.x) + fo + iy * p.Wraw));
          }
          vmask |= 1u << r;
        }
      }
    };

    StemCursor cur(blockIdx.x, gridDim.x, p.tiles_per_frame);
    for (int i = 0; i < lw; ++i) cur.advance();
    cur.ft = cur.frame % p.T;
    if (cur.frame < p.frames) issue(cur);
    for (uint32_t it = 0; cur.frame < p.frames; ++it) {
      const int slot = lw + kLoaders * (it & 1);            // stage s = lw + 3*it  ->  slot s % 6
      __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(strip + slot * strip_buf);
      mbar_wait(&sempty[slot], ((it >> 1) & 1) ^ 1);
      if (grp_ok) {
#pragma unroll
        for (int r = 0; r < kIters; ++r) {
          if (r < p.strip_rows) {
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            if (vmask & (1u << r)) {
              if constexpr (kU8) {
                const uint32_t u = raw[r];
                v0 = fmaf((float)(u & 0xffu), p.u8_scale, p.u8_bias);
                v1 = fmaf((float)((u >> 8) & 0xffu), p.u8_scale, p.u8_bias);
                v2 = fmaf((float)((u >> 16) & 0xffu), p.u8_scale, p.u8_bias);
                v3 = fmaf((float)(u >> 24), p.u8_scale, p.u8_bias);
              } else {
                v0 = raw[r].x; v1 = raw[r].y; v2 = raw[r].z; v3 = raw[r].w;
              }
            }
            __nv_bfloat16* row = sb + r * SP + 4 * grp;   // strip column c = ix + 3 = 4*grp - 1 + j
            if (grp > 0) row[-1] = __float2bfloat16_rn(v0);
            *reinterpret_cast<uint32_t*>(row) = pack_bf16x2(v1, v2);
            row[2] = __float2bfloat16_rn(v3);
          }
        }
      }
      const int prev_frame = cur.frame;
      cur.advance(); cur.advance(); cur.advance();
      if (cur.frame != prev_frame) cur.ft = cur.frame % p.T;
      if (cur.frame < p.frames) issue(cur);                 // next own stage's loads fly while builders work
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[slot]);
    }
  } else if (warp == 4) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      mbar_expect_tx(wbar, kStemBBytes);
      for (int kt = 0; kt < 5; ++kt) tma_load_2d(smB + kt * 8192, &mapW, wbar, kt * 64, 0);
      mbar_wait(wbar, 0);
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int frame = blockIdx.x; frame < p.frames; frame += gridDim.x) {
        for (int tile = 0; tile < p.tiles_per_frame; ++tile) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 64;
          for (int kt = 0; kt < 5; ++kt) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smA + stage * kStemABytes));
            const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smB + kt * 8192));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (kt | k) != 0 ? 1u : 0u);
            umma_commit(&empty[stage]);
            if (++stage == kStemAStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // =============================================================== epilogue: BN + PReLU -> ring -> max-pool
    const int et = threadIdx.x;                 // 0..127 ; warp == TMEM lane quarter
    int acc = 0;
    uint32_t acc_phase = 0;
    const int ring_mask = p.ring_rows - 1;
    const int row_bytes = p.Wo * 128;
    for (int frame = blockIdx.x; frame < p.frames; frame += gridDim.x) {
      int py_done = 0;
      uint16_t* yframe = p.y + (size_t)frame * p.out_img_rows * p.Wp * 64;
      for (int tile = 0; tile < p.tiles_per_frame; ++tile) {
        const int m = tile * 128 + et;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64;
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tempty[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        if (m < p.Mf) {
          const int yy = m / p.Wo, xx = m - yy * p.Wo;
          uint8_t* dst = ring + (yy & ring_mask) * row_bytes + xx * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = ch * 8 + i;
              const float a = __uint_as_float(c < 32 ? r0[c] : r1[c - 32]);
              const float z = fmaf(a, chan[c], chan[64 + c]);
              v[i] = z > 0.f ? z : z * chan[128 + c];
            }
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(dst + ((ch ^ (xx & 7)) << 4)) = o;
          }
        }
        named_bar_sync(2, 128);
        // pooled rows whose three conv rows are now complete
        const int m_end = min((tile + 1) * 128, p.Mf);
        const int rows_complete = m_end / p.Wo;               // conv rows 0 .. rows_complete-1 are final
        const int py_ready = rows_complete / 2;               // needs conv row 2*py+1 <= rows_complete-1
        const int items = (py_ready - py_done) * p.Wp * 8;
        for (int it = et; it < items; it += 128) {
          const int ch = it & 7;
          const int pix = it >> 3;
          const int pyo = pix / p.Wp, px = pix - pyo * p.Wp;
          const int py = py_done + pyo;
          __nv_bfloat162 best[4];
          bool first = true;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int cy = 2 * py + dy;
            if (cy < 0) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int cx = 2 * px + dx;
              if (cx < 0) continue;
              const uint4 v = *reinterpret_cast<const uint4*>(ring + (cy & ring_mask) * row_bytes + cx * 128 +
                                                              ((ch ^ (cx & 7)) << 4));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
              if (first) {
                best[0] = h[0]; best[1] = h[1]; best[2] = h[2]; best[3] = h[3];
                first = false;
              } else {
                best[0] = __hmax2(best[0], h[0]); best[1] = __hmax2(best[1], h[1]);
                best[2] = __hmax2(best[2], h[2]); best[3] = __hmax2(best[3], h[3]);
              }
            }
          }
          *reinterpret_cast<uint4*>(yframe + ((size_t)py * p.Wp + px) * 64 + ch * 8) =
              *reinterpret_cast<const uint4*>(best);
        }
        py_done = py_ready;
        named_bar_sync(2, 128);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace dl

extern "C" int dl_stem_conv3d_bn_prelu_pool(const void* x, int is_u8, int B, int T, int H, int W, int Hraw, int Wraw,
                                            float mean, float std, const void* w_packed, const float* scale,
                                            const float* shift, const float* slope, void* y, int out_img_rows,
                                            void* stream) {
  using namespace dl;
  DL_CHECK_ARG(x && w_packed && scale && shift && slope && y, "stem: null pointer");
  DL_CHECK_ARG(B > 0 && T > 0, "stem: empty batch");
  DL_CHECK_ARG(H >= 8 && W >= 32 && H % 4 == 0 && W % 4 == 0 && W <= 116, "stem: H, W must be multiples of 4, 32 <= W <= 116");
  if (is_u8) {
    DL_CHECK_ARG(Hraw >= H && Wraw >= W && std != 0.f, "stem: raw crop smaller than the centre crop");
  }
  int st = require_sm100();
  if (st != DL_OK) return st;

  StemParams p;
  p.x = x;
  p.B = B; p.T = T; p.H = H; p.W = W; p.Hraw = Hraw; p.Wraw = Wraw;
  // CenterCrop: delta = int(round(w - tw) / 2.)  (models/video_models/preprocess.py:88-90)
  p.dh = is_u8 ? (Hraw - H) / 2 : 0;
  p.dw = is_u8 ? (Wraw - W) / 2 : 0;
  p.u8_scale = is_u8 ? 1.0f / (255.0f * std) : 1.0f;
  p.u8_bias = is_u8 ? -mean / std : 0.0f;
  if (is_u8) {
    DL_CHECK_ARG(Wraw % 4 == 0 && p.dw % 4 == 0, "stem: u8 crops need Wraw and the crop offset to be multiples of 4");
  }
  p.Ho = H / 2; p.Wo = W / 2; p.Hp = H / 4; p.Wp = W / 4;
  p.Mf = p.Ho * p.Wo;
  p.tiles_per_frame = (p.Mf + 127) / 128;
  const int span = (127 + p.Wo - 1) / p.Wo;           // extra conv rows a tile can reach past its first row
  int rr = 1;
  while (rr < span + 6) rr <<= 1;
  p.ring_rows = rr;
  p.strip_rows = 2 * span + 7;
  p.strip_pitch = W + 12;
  DL_CHECK_ARG(p.strip_rows <= kStripRowsMax && W / 4 + 2 < 32, "stem: needs 32 <= W <= 116");
  p.scale = scale; p.shift = shift; p.slope = slope;
  p.y = static_cast<uint16_t*>(y);
  p.frames = B * T;
  p.out_img_rows = out_img_rows > 0 ? out_img_rows : p.Hp;
  DL_CHECK_ARG(p.out_img_rows >= p.Hp, "stem: out_img_rows < H/4");

  const int strip_elems = p.strip_rows * p.strip_pitch;
  const size_t smem = 1024 + (size_t)kStemAStages * kStemABytes + kStemBBytes + (size_t)p.ring_rows * p.Wo * 128 +
                      kStripSlots * (size_t)((strip_elems + 7) & ~7) * 2 + 192 * 4 + 32 * 8 + 16 + 64 * 4;
  DL_CHECK_ARG(p.tiles_per_frame <= 64, "stem: frame too large (more than 64 tiles)");
  DL_CHECK_ARG(smem <= 227 * 1024, "stem: shared-memory budget exceeded (%zu B)", smem);
  CUtensorMap mapW;
  st = make_tiled_2d_bf16(&mapW, w_packed, 64, 320, 320, 64, 64);
  if (st != DL_OK) return st;
  int grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (p.frames < grid) grid = p.frames;
  const bool small = p.strip_rows <= 13;      // strip rows per stage: 13 (GRID 88x88) or up to 24
  void (*kern)(const CUtensorMap, const StemParams) =
      is_u8 ? (small ? stem_conv3d_kernel<true, 13> : stem_conv3d_kernel<true, 24>)
            : (small ? stem_conv3d_kernel<false, 13> : stem_conv3d_kernel<false, 24>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(DL_ERR_CUDA, "stem smem attribute: %s", cudaGetErrorString(e));
  kern<<<grid, kStemThreads, smem, (cudaStream_t)stream>>>(mapW, p);
  return check_launch("stem_conv3d_kernel");
}
