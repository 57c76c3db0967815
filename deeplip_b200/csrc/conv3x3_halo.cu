// 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels (ResNet layer1, models/video_models/resnet.py:56-69) with
// operand reuse in shared memory.
//
// The generic implicit-GEMM kernel (igemm_conv.cu) fetches one [128 px x 64 ch] im2col box per filter tap, i.e.
// every activation crosses L2 -> SM nine times; with only 64 output channels per tile that makes layer1
// L2-bandwidth bound (measured ~10 TB/s, 445 TFLOP/s).  Here ONE TMA box -- an 18-row x 10-pixel halo patch --
// feeds all nine taps: the A operand of tap (r, s) is simply a shifted view of the patch, expressed through the
// UMMA shared-memory descriptor (start + (10 r + s) * 128 B, 1280 B between 8-pixel row groups), and the whole
// 64 x 576 weight matrix stays resident in shared memory.  L2 traffic per output tile drops from 216 KB to 23 KB.
//
// Layout contract ("stacked rows"): activations are (N, img_rows, W, 64) bf16 with img_rows >= H + 1 and rows
// H .. img_rows-1 of every image all zero.  Stacked row R = n * img_rows + y; the zero row is at once the bottom
// padding of image n and the top padding of image n + 1, so a tile may span images.  Left / right padding comes
// from TMA out-of-bounds zero fill.  An output tile is 16 stacked rows x 8 columns (M = 128); column groups of a
// row start at 0, 8, ..., W - 8 (the last one may overlap its neighbour: those pixels are written twice with identical
// values).  The padding rows of y are written as zeros, so y can be used as an input again.
#include "dl_host.cuh"
#include "dl_ptx.cuh"

namespace dl {

constexpr int kHaloPatchRows = 18, kHaloPatchCols = 10;                  // 16 x 8 output pixels + a one-pixel halo
constexpr int kHaloPatchBytes = kHaloPatchRows * kHaloPatchCols * 128;   // 23 040 bytes landed by TMA
constexpr int kHaloPatchStride = (kHaloPatchBytes + 1023) / 1024 * 1024; // stage stride: patch bases stay 1024-byte aligned
constexpr int kHaloStages = 5;
constexpr int kHaloWBytes = 9 * 64 * 64 * 2;                             // 73 728 resident weights
constexpr int kHaloThreads = 576;                                        // TMA + MMA warps, 2 epilogue groups x 8 warps
constexpr int kHaloAccs = 4;                                             // TMEM accumulators (64 columns each)
constexpr int kHaloOutBytes = 128 * 128;                                 // one output tile (128 px x 64 ch bf16) staged for the TMA store
constexpr int kHaloSmem = kHaloStages * kHaloPatchStride + kHaloWBytes + 2 * kHaloOutBytes + 3 * 64 * 4 + 24 * 8 + 16 + 1024;

struct HaloParams {
  int rows_total, img_rows, H, W;
  int num_groups, total_tiles;
  const float* scale;
  const float* shift;
  const float* slope;
  const uint16_t* residual;
  uint16_t* y;
};

__global__ void __launch_bounds__(kHaloThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
                    const __grid_constant__ CUtensorMap mapY, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* patches = smem;
  uint8_t* wres = smem + kHaloStages * kHaloPatchStride;
  uint8_t* outst = wres + kHaloWBytes;                      // 2 x output tile, one per epilogue group (1024-aligned)
  float* prm = reinterpret_cast<float*>(outst + 2 * kHaloOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(prm + 192);
  uint64_t* full = bars;
  uint64_t* empty = bars + kHaloStages;
  uint64_t* tfull = bars + 2 * kHaloStages;
  uint64_t* tempty = tfull + kHaloAccs;
  uint64_t* wfull = tempty + kHaloAccs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kHaloStages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int a = 0; a < kHaloAccs; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], 256);          // the 8 warps of the epilogue group that owns the accumulator
      }
      mbar_init(wfull, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<64 * kHaloAccs>(tmem_slot);
  }
  for (int c = threadIdx.x; c < 64; c += kHaloThreads) {
    prm[c] = p.scale[c];
    prm[64 + c] = p.shift[c];
    prm[128 + c] = p.slope[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(wfull, kHaloWBytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(wres + tap * 8192, &mapW, wfull, tap * 64, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int rt = tile / p.num_groups;
        const int g = tile - rt * p.num_groups;
        const int x0 = min(8 * g, p.W - 8);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], kHaloPatchBytes);
        tma_load_3d(patches + stage * kHaloPatchStride, &mapX, &full[stage], 0, x0 - 1, rt * 16 - 1);
        if (++stage == kHaloStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {           // one thread, known to ptxas as such: no per-lane loops around UTCHMMA
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      mbar_wait(wfull, 0);
      const uint32_t wbase = smem_u32(wres);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 64;
        const uint32_t pbase = smem_u32(patches + stage * kHaloPatchStride);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - 3 * r;
          // rows of the A view: output pixel (rr, xx) reads patch pixel (rr + r, xx + s); pitch 10 px = 1280 B
          const uint32_t a0 = pbase + (uint32_t)(r * kHaloPatchCols + s) * 128u;
          // The 128B swizzle is a function of the absolute shared-memory address bits (verified on B200: a view that
          // starts s pixels into the row needs NO descriptor base offset; setting one scrambles the operand).
          const uint32_t bo = 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = umma_desc_sw128_kmajor_ex(a0 + 32u * k, kHaloPatchCols * 128u, bo);
            const uint64_t bdesc = umma_desc_sw128_kmajor(wbase + tap * 8192u + 32u * k);
            umma_bf16(d, adesc, bdesc, idesc, (tap | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == kHaloStages) { stage = 0; phase ^= 1; }
        if (++acc == kHaloAccs) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // Two epilogue groups of 8 warps take alternate tiles (group g: tiles g, g+2, ... of this CTA; accumulator
    // i & 3), so one group's global-memory round trips (residual loads, stores) overlap the other's and the MMA
    // thread can run two tiles ahead.
    const int quarter = warp & 3;
    const int chunk = ((warp - 2) >> 2) & 1;    // 32-channel half handled by this warp
    const int egroup = (warp - 2) >> 3;
    const bool has_res = p.residual != nullptr;
    const int m = quarter * 32 + lane;          // accumulator row == TMEM lane
    const int rr = m >> 3, xx = m & 7;
    int acc = egroup;
    uint32_t acc_phase = 0;
    // Output tiles leave through shared memory and ONE TMA store per tile (full 128-byte lines, asynchronous) instead
    // of 16-byte stores at a 128-byte stride.  The box covers all 16 x 8 pixels: pixels of the zero rows between
    // images are staged as zeros (so they stay zero), rows past the tensor end are clipped by the TMA unit, and the
    // columns the last column group shares with its neighbour are written twice with identical values.
    uint8_t* stg = outst + egroup * kHaloOutBytes;
    const bool issuer = (warp & 7) == 2 && lane == 0;       // first warp of each group (warps 2 and 10)
    for (int tile = blockIdx.x + egroup * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x) {
      const int rt = tile / p.num_groups;
      const int g = tile - rt * p.num_groups;
      const int x0 = min(8 * g, p.W - 8);
      const int R = rt * 16 + rr;
      const bool inimg = R < p.rows_total && (R % p.img_rows) < p.H;
      // residual straight from global memory into registers, requested before the wait on the accumulator
      uint4 res[4];
      if (has_res && inimg) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + ((size_t)R * p.W + x0 + xx) * 64 + chunk * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) res[q] = __ldg(rp + q);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) res[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      uint32_t acc_r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 64 + chunk * 32, acc_r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc += 2;
      if (acc >= kHaloAccs) { acc -= kHaloAccs; acc_phase ^= 1; }
      const float4* sc = reinterpret_cast<const float4*>(prm + chunk * 32);
      const float4* sh = reinterpret_cast<const float4*>(prm + 64 + chunk * 32);
      const float4* sl = reinterpret_cast<const float4*>(prm + 128 + chunk * 32);
      uint4 o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 s4 = sc[2 * q + h], h4 = sh[2 * q + h];
          v[4 * h + 0] = fmaf(__uint_as_float(acc_r[8 * q + 4 * h + 0]), s4.x, h4.x);
          v[4 * h + 1] = fmaf(__uint_as_float(acc_r[8 * q + 4 * h + 1]), s4.y, h4.y);
          v[4 * h + 2] = fmaf(__uint_as_float(acc_r[8 * q + 4 * h + 2]), s4.z, h4.z);
          v[4 * h + 3] = fmaf(__uint_as_float(acc_r[8 * q + 4 * h + 3]), s4.w, h4.w);
        }
        if (has_res) {
          const uint4 rv = res[q];
          v[0] += bf16_lo(rv.x); v[1] += bf16_hi(rv.x); v[2] += bf16_lo(rv.y); v[3] += bf16_hi(rv.y);
          v[4] += bf16_lo(rv.z); v[5] += bf16_hi(rv.z); v[6] += bf16_lo(rv.w); v[7] += bf16_hi(rv.w);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 l4 = sl[2 * q + h];
          v[4 * h + 0] = v[4 * h + 0] > 0.f ? v[4 * h + 0] : v[4 * h + 0] * l4.x;
          v[4 * h + 1] = v[4 * h + 1] > 0.f ? v[4 * h + 1] : v[4 * h + 1] * l4.y;
          v[4 * h + 2] = v[4 * h + 2] > 0.f ? v[4 * h + 2] : v[4 * h + 2] * l4.z;
          v[4 * h + 3] = v[4 * h + 3] > 0.f ? v[4 * h + 3] : v[4 * h + 3] * l4.w;
        }
        o[q].x = pack_bf16x2(v[0], v[1]); o[q].y = pack_bf16x2(v[2], v[3]);
        o[q].z = pack_bf16x2(v[4], v[5]); o[q].w = pack_bf16x2(v[6], v[7]);
      }
      // the group's previous TMA store must have finished reading the staging tile
      if (issuer) bulk_wait_group_read0();
      named_bar_sync(3 + egroup, 256);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(stg + m * 128 + (((chunk * 4 + q) ^ (m & 7)) << 4)) = inimg ? o[q] : make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();                   // generic-proxy writes -> visible to the TMA unit
      named_bar_sync(3 + egroup, 256);
      if (issuer) {
        tma_store_3d(&mapY, stg, 0, x0, rt * 16);
        bulk_commit_group();
      }
    }
  }

  if (warp >= 2 && (warp & 7) == 2 && lane == 0) bulk_wait_group0();     // outstanding output stores
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64 * kHaloAccs>(tmem_base);
  }
}

}  // namespace dl

extern "C" int dl_conv3x3_c64_halo_bf16(const void* x, const void* w_packed, const float* scale, const float* shift,
                                        const float* slope, const void* residual, void* y, int N, int H, int W,
                                        int img_rows, void* stream) {
  using namespace dl;
  DL_CHECK_ARG(x && w_packed && scale && shift && slope && y, "conv3x3_halo: null pointer");
  DL_CHECK_ARG(N > 0 && H > 0 && W >= 8 && img_rows >= H + 1, "conv3x3_halo: need W >= 8 and img_rows >= H + 1");
  DL_CHECK_ARG((long long)N * img_rows < (1ll << 31) / 16, "conv3x3_halo: too many rows");
  int st = require_sm100();
  if (st != DL_OK) return st;
  static PerDevice<bool> configured_dev;
  bool* configured = configured_dev.slot();
  if (!configured) return fail(DL_ERR_CUDA, "conv3x3_halo: no current device");
  if (!*configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmem);
    if (e != cudaSuccess) return fail(DL_ERR_CUDA, "conv3x3_halo smem attribute: %s", cudaGetErrorString(e));
    *configured = true;
  }
  HaloParams p;
  p.rows_total = N * img_rows; p.img_rows = img_rows; p.H = H; p.W = W;
  p.num_groups = (W + 7) / 8;
  p.total_tiles = ((p.rows_total + 15) / 16) * p.num_groups;
  p.scale = scale; p.shift = shift; p.slope = slope;
  p.residual = static_cast<const uint16_t*>(residual);
  p.y = static_cast<uint16_t*>(y);
  CUtensorMap mapX, mapW, mapY;
  st = make_tiled_3d_bf16(&mapX, x, (uint64_t)p.rows_total, (uint64_t)W, 64, kHaloPatchRows, kHaloPatchCols, 64);
  if (st != DL_OK) return st;
  st = make_tiled_2d_bf16(&mapW, w_packed, 64, 576, 576, 64, 64);
  if (st != DL_OK) return st;
  st = make_tiled_3d_bf16(&mapY, y, (uint64_t)p.rows_total, (uint64_t)W, 64, 16, 8, 64);
  if (st != DL_OK) return st;
  int grid = device_sm_count();
  if (grid <= 0) grid = 148;
  if (p.total_tiles < grid) grid = p.total_tiles;
  conv3x3_halo_kernel<<<grid, kHaloThreads, kHaloSmem, (cudaStream_t)stream>>>(mapX, mapW, mapY, p);
  return check_launch("conv3x3_halo_kernel");
}
