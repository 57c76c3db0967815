// Shared pieces of the implicit-GEMM kernels (1-CTA igemm_conv.cu and CTA-pair igemm2_conv.cu): kernel
// parameters and the fused epilogue (folded BN scale/shift, residual add, per-channel slope, bf16 store and the
// optional fp32 side output), operating on one 128-row accumulator tile held in this CTA's tensor memory.
#pragma once
#include "dl_host.cuh"
#include "dl_ptx.cuh"

namespace dl {

constexpr int kMaxCout = 2048;   // per-channel epilogue parameters staged in shared memory

struct IgemmParams {
  int M, P, Q, PQ;
  int Cout;
  int cchunks, R, S;
  int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int num_m_blocks, num_n_blocks;
  int ldy, ldf;
  int lin, lin_w, lin_h, valid_w, valid_h;   // guarded-linear operand A (tiled TMA): geometry and stored extents
  int dbg;
  int taps, a_rows;                          // pair kernel: horizontal taps sharing one A box of a_rows rows
  int out_hp, out_wp;                        // im2col mode writing into a guarded tensor (0 = dense)
  float f32_slope;
  const float* scale;
  const float* shift;
  const float* slope;
  const float* scale2;
  const float* shift2;
  const uint16_t* residual;
  uint16_t* y;
  float* yf;
  uint16_t* y2;        // split output: channels >= split_c go to y2[row * ldy + c - split_c] (split_c = 0: off)
  int split_c;
  int skip_n0;         // pair kernel: n blocks >= skip_n0 see the filter's centre tap only (-1: none)
  int half_skip;       // pair kernel, one resident 256-wide n block: channels >= 128 see the centre tap only -> two N=128 MMA streams
  // K4 in the epilogue (pair kernel, staged 256-wide streaming variant): every pool_rows consecutive GEMM rows are one
  // image; their per-channel mean goes to pool_out[image * Cout + c].  Tiles then step by tile_rows = the largest
  // multiple of pool_rows <= 128 rows instead of 128, so that no image straddles two tiles (the MMA still covers 128
  // rows: the surplus rows are the next tile's first rows, computed twice and pooled once).
  int tile_rows;       // GEMM rows between consecutive m tiles (128 unless pooling)
  int pool_rows;       // 0 = off
  int store_y;         // 0: pooling mode without the bf16 output tensor
  float* pool_out;
};

// GEMM row -> row of y / residual, and whether it is stored at all.
__device__ __forceinline__ bool igemm_map_row(const IgemmParams& p, long long& row) {
  if (row >= p.M) return false;
  if (p.lin) {                                   // same geometry in and out; guard positions are never written
    const int r = (int)row;
    const int line = r / p.lin_w;
    return r - line * p.lin_w < p.valid_w && line % p.lin_h < p.valid_h;
  }
  if (p.out_wp > 0) {
    const int r = (int)row;
    const int img = r / p.PQ;
    const int rem = r - img * p.PQ;
    const int pp = rem / p.Q;
    row = ((long long)img * p.out_hp + pp) * p.out_wp + (rem - pp * p.Q);
  }
  return true;
}

// Residual rows of this warp's first 32-channel chunk, requested BEFORE the wait on the accumulator so that the
// L2 round trip overlaps the tile's MMAs.
__device__ __forceinline__ void igemm_prefetch_residual(const IgemmParams& p, long long row, bool row_ok, int cbase,
                                                        int chunk0, bool has_res, uint4 (&res)[4]) {
  const int c0 = cbase + chunk0 * 32;
  if (has_res && row_ok && c0 + 32 <= p.Cout) {
    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row * p.ldy + c0);
#pragma unroll
    for (int g = 0; g < 4; ++g) res[g] = __ldg(rp + g);
  }
}

// Epilogue of one tile for one warp: TMEM lane quarter `quarter`, 32-column chunks chunk0, chunk0+2, ...
// tmem_acc = TMEM address of the accumulator stage (column of channel cbase, lane 0); prm = scale | shift | slope
// arrays in shared memory, pstride floats apart.
template <int BLOCK_N>
__device__ __forceinline__ void igemm_epilogue_tile(const IgemmParams& p, const float* prm, int pstride,
                                                    uint32_t tmem_acc, long long row, bool row_ok, int cbase,
                                                    int quarter, int chunk0, bool has_res, bool fast,
                                                    uint4 (&res)[4]) {
  // split output (sibling convs in one launch): a tile lies on one side of the split (split_c % BLOCK_N == 0)
  uint16_t* const ybase = (p.split_c > 0 && cbase >= p.split_c) ? p.y2 - p.split_c : p.y;
#pragma unroll 1
  for (int j = chunk0; j < BLOCK_N / 32; j += 2) {
    const int c0 = cbase + j * 32;
    if (c0 >= p.Cout) break;
    uint32_t acc_r[32];
    tmem_ld_32x32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + j * 32, acc_r);
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
      const float4* sh = reinterpret_cast<const float4*>(prm + pstride + c0);
      const float4* sl = reinterpret_cast<const float4*>(prm + 2 * pstride + c0);
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
        uint4* dst = reinterpret_cast<uint4*>(ybase + row * p.ldy + c0);
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
            for (int i = 0; i < 8; ++i) v[i] = fmaf(a[i], prm[c + i], prm[pstride + c + i]);
            if (has_res) {
              const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ldy + c));
              v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
              v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * prm[2 * pstride + c + i];
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(ybase + row * p.ldy + c) = o;
          }
        }
      }
    }
  }
}

// Staged form of the fast path above for the CTA-pair kernels (dense row-major y; channels past Cout in the last slab
// are computed on garbage and clipped by the tensor map):
// the 8 epilogue warps of a CTA assemble one [128 rows x 64 channels] bf16 slab at a time in a 16 KB shared-memory
// buffer (128-byte rows, 128B swizzle -- the layout the TMA store expects) and one thread sends it with ONE
// cp.async.bulk.tensor store: full 128-byte lines, asynchronous, rows past M clipped by the tensor map.  NBUF = 2:
// slab k+1 is assembled while slab k drains (one named barrier per slab); NBUF = 1: the buffer is reclaimed first
// (two barriers per slab).  Warps with chunk0 = 0 / 1 fill the lower / upper 32 channels of every row.
// Arithmetic and rounding are those of igemm_epilogue_tile: same bits.
template <int BLOCK_N, int NBUF>
__device__ __forceinline__ void igemm_epilogue_tile_staged(const IgemmParams& p, const float* prm, int pstride,
                                                           uint32_t tmem_acc, long long row, bool row_ok, int tile_row0,
                                                           int cbase, int quarter, int chunk0, bool has_res,
                                                           uint4 (&res)[4], uint8_t* stg, int& sbuf, const void* mapY,
                                                           int col0, bool issuer) {
  const int m = quarter * 32 + (int)lane_id();
#pragma unroll 1
  for (int i = 0; i < BLOCK_N / 64; ++i) {
    if (cbase + 64 * i >= p.Cout) break;                 // the same for all eight warps
    const int j = 2 * i + chunk0;
    const int c0 = cbase + j * 32;
    uint32_t acc_r[32];
    tmem_ld_32x32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + j * 32, acc_r);
    uint4 res_cur[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) res_cur[g] = res[g];
    if (has_res && row_ok && i + 1 < BLOCK_N / 64 && c0 + 96 <= p.Cout) {      // prefetch the next slab's residual
      const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row * p.ldy + c0 + 64);
#pragma unroll
      for (int g = 0; g < 4; ++g) res[g] = __ldg(rp + g);
    }
    tmem_ld_wait();
    const float4* sc = reinterpret_cast<const float4*>(prm + c0);
    const float4* sh = reinterpret_cast<const float4*>(prm + pstride + c0);
    const float4* sl = reinterpret_cast<const float4*>(prm + 2 * pstride + c0);
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
      if (has_res && row_ok) {
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
    uint8_t* sb = stg + sbuf * (128 * 128);
    if (NBUF == 1) {                                     // reclaim the only buffer: its last store has read it
      if (issuer) bulk_wait_group_read0();
      named_bar_sync(1, 256);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(sb + m * 128 + (((chunk0 * 4 + g) ^ (m & 7)) << 4)) = o[g];
    fence_proxy_async_smem();                            // generic-proxy writes -> visible to the TMA unit
    // NBUF == 2: the previous slab's store (other buffer) must have read its buffer before anyone passes this barrier
    // and starts on the slab after this one
    if (NBUF == 2 && issuer) bulk_wait_group_read0();
    named_bar_sync(1, 256);
    if (issuer && !(p.dbg & 2) && p.store_y) {
      tma_store_2d(mapY, sb, col0 + 64 * i, tile_row0);      // col0: this n block's first channel in mapY's tensor
      bulk_commit_group();
    }
    if (NBUF == 2 && p.pool_rows > 0) {
      // ---- K4: per-image mean of this slab's 64 channels, read back from the staged bf16 rows (the values the output
      // tensor holds).  Thread = (image f of the tile, 4 channels); the pool_rows rows are added in order from 0.f and
      // scaled by 1 / pool_rows: the arithmetic of frame_pool_kernel, bit for bit.  The slab buffer is rewritten two
      // slabs later, behind the next slab's barrier, which every thread reaches only after this pass.
      const int et = (int)threadIdx.x - 64;                  // epilogue warps are warps 2..9
      const int f = et >> 4, q = et & 15;
      if (f * p.pool_rows < p.tile_rows) {
        const long long img = (long long)(tile_row0 / p.pool_rows) + f;
        if ((img + 1) * p.pool_rows <= (long long)p.M) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          for (int r = 0; r < p.pool_rows; ++r) {
            const int mm = f * p.pool_rows + r;
            const uint2 v = *reinterpret_cast<const uint2*>(sb + mm * 128 + (((q >> 1) ^ (mm & 7)) << 4) + (q & 1) * 8);
            s0 += bf16_lo(v.x); s1 += bf16_hi(v.x); s2 += bf16_lo(v.y); s3 += bf16_hi(v.y);
          }
          const float inv = 1.f / (float)p.pool_rows;
          *reinterpret_cast<float4*>(p.pool_out + img * p.Cout + cbase + 64 * i + 4 * q) =
              make_float4(s0 * inv, s1 * inv, s2 * inv, s3 * inv);
        }
      }
    }
    if (NBUF == 2) sbuf ^= 1;
  }
}

}  // namespace dl
