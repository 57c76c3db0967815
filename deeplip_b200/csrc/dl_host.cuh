// Host-side glue shared by every translation unit of libdeeplip_b200.so:
// thread-local error string, launch counter, driver-API tensor-map encoders.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/deeplip_b200.h"

namespace dl {

int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);          // cudaGetLastError -> DL_OK / DL_ERR_CUDA
int current_device();      // cudaGetDevice, or -1
int device_sm_count();     // of the current device
int require_sm100();       // of the current device

// Once-per-device host state (function attributes, constant tables): a process may drive several GPUs
// (the reference's nn.DataParallel usage, a manual cuda:1 call), and each device needs its own setup.
constexpr int kMaxDevices = 64;
template <typename T>
struct PerDevice {
  T v[kMaxDevices] = {};
  T* slot() { const int d = current_device(); return (d >= 0 && d < kMaxDevices) ? &v[d] : nullptr; }
};
int opt_pair();            // tuning switches (dl_set_option): CTA-pair kernels on / off
int opt_dbg();
int opt_frontend();        // 2 = register-resident radix-8 FFT front end (default), 1 = first-generation kernels
int opt_prepass();         // 2 = one block per frame, aligned word loads (default), 1 = first-generation stem pre-pass
int opt_small_linear();    // 1 = fc layers (P Q == 1, <= 4096 rows) run on linear_small_kernel (default), 0 = igemm
int opt_statpool_mlp();    // stat pool: 16-byte loads in flight per lane, 4 (default; measured faster) or 8
int opt_statpool_slab();   // stat pool: channels per block, 256 (default) or 128 (half a warp per time step; measured slower)
int opt_stem();            // 2 = channels-on-lanes stem kernel for W = 88 (default), 1 = first-generation kernel for every shape
int opt_pool_fuse();       // 1 = dl_conv_desc.avgpool is taken in the pair kernel's epilogue where the shape allows it (default), 0 = always the pooling kernel
int opt_stft_pad();        // stft centre padding: 0 reflect (librosa < 0.10), 1 zeros (librosa >= 0.10)
int opt_tap_share();     // pair kernel shares one operand-A box across horizontal taps (guarded-linear mode)
int opt_staged_epilogue();   // 1 = resident pair kernels store their tiles through shared memory + TMA (default), 0 = per-lane stores
int opt_pair_resident();   // resident weight-half variant of the pair kernel on / off                          // DL_OK or DL_ERR_UNSUPPORTED

// 2-D tiled map over a row-major (rows, cols) 16-bit matrix with row pitch `ld` elements;
// box = (box_cols, box_rows), 128-byte swizzle.
int make_tiled_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols);

// 3-D tiled map over a (rows, cols, C) 16-bit tensor (C contiguous): box = (box_c, box_cols, box_rows).
int make_tiled_3d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t C,
                       uint32_t box_rows, uint32_t box_cols, uint32_t box_c);

// 4-D tiled map, no swizzle, over a dense (d3, d2, rows, cols) 16-bit tensor; box = (box_cols, box_rows, box_d2, 1).
int make_tiled_4d_bf16_noswizzle(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t d2,
                                 uint64_t d3, uint32_t box_cols, uint32_t box_rows, uint32_t box_d2 = 1);

// im2col map over an NHWC 16-bit activation tensor (pitch ldx elements per pixel): loads
// `pixels` output positions x `channels` channels per request, 128-byte swizzle, zero OOB fill.
int make_im2col_nhwc_bf16(CUtensorMap* map, const void* base, int N, int H, int W, int C, int ldx, int img_rows,
                          int img_cols, int R, int S,
                          int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                          uint32_t channels, uint32_t pixels);

}  // namespace dl

#define DL_CHECK_ARG(cond, ...) \
  do { if (!(cond)) return dl::fail(DL_ERR_INVALID, __VA_ARGS__); } while (0)
