#include "dl_host.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace dl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DL_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  count_launch();
  return DL_OK;
}

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int device_sm_count() {
  static PerDevice<int> sms;
  int* s = sms.slot();
  if (!s) return 0;
  if (*s == 0) cudaDeviceGetAttribute(s, cudaDevAttrMultiProcessorCount, current_device());
  return *s;
}

int require_sm100() {
  static PerDevice<int> major;          // 0 = not queried yet
  int* m = major.slot();
  if (!m) return fail(DL_ERR_CUDA, "no CUDA device (cudaGetDevice failed or device index >= %d)", kMaxDevices);
  if (*m == 0) cudaDeviceGetAttribute(m, cudaDevAttrComputeCapabilityMajor, current_device());
  if (*m != 10) return fail(DL_ERR_UNSUPPORTED, "deeplip_b200 needs an sm_100 device (found sm_%d0)", *m);
  return DL_OK;
}

// ---------------------------------------------------------------- driver entry points
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_tiled = nullptr;
static EncodeIm2colFn g_im2col = nullptr;
static int g_driver_version = 0;

static int load_driver() {
  static std::once_flag once;
  static int status = DL_OK;
  std::call_once(once, [] {
    cudaFree(0);  // make sure a context exists
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      status = DL_ERR_CUDA;
      return;
    }
    g_tiled = reinterpret_cast<EncodeTiledFn>(f);
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      status = DL_ERR_CUDA;
      return;
    }
    g_im2col = reinterpret_cast<EncodeIm2colFn>(f);
    cudaDriverGetVersion(&g_driver_version);
  });
  if (status != DL_OK) return fail(DL_ERR_CUDA, "cuTensorMapEncode* driver entry points unavailable");
  return DL_OK;
}

int make_tiled_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols) {
  int st = load_driver();
  if (st != DL_OK) return st;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return DL_OK;
}

int make_tiled_3d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t C,
                       uint32_t box_rows, uint32_t box_cols, uint32_t box_c) {
  int st = load_driver();
  if (st != DL_OK) return st;
  cuuint64_t dims[3] = {C, cols, rows};
  cuuint64_t strides[2] = {C * 2, cols * C * 2};
  cuuint32_t box[3] = {box_c, box_cols, box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DL_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d): rows=%llu cols=%llu C=%llu", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)C);
  return DL_OK;
}

int make_tiled_4d_bf16_noswizzle(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t d2,
                                 uint64_t d3, uint32_t box_cols, uint32_t box_rows, uint32_t box_d2) {
  int st = load_driver();
  if (st != DL_OK) return st;
  cuuint64_t dims[4] = {cols, rows, d2, d3};
  cuuint64_t strides[3] = {cols * 2, rows * cols * 2, d2 * rows * cols * 2};
  cuuint32_t box[4] = {box_cols, box_rows, box_d2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DL_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed (%d): cols=%llu rows=%llu d2=%llu d3=%llu box=%ux%u",
                (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)d2,
                (unsigned long long)d3, box_cols, box_rows);
  return DL_OK;
}

int make_im2col_nhwc_bf16(CUtensorMap* map, const void* base, int N, int H, int W, int C, int ldx, int img_rows,
                          int img_cols, int R, int S,
                          int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                          uint32_t channels, uint32_t pixels) {
  int st = load_driver();
  if (st != DL_OK) return st;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)img_cols * ldx * 2,
                           (cuuint64_t)img_rows * img_cols * ldx * 2};
  // Bounding box of the filter's top-left anchor: starts at -pad and stops so that the last tap
  // (offset (S-1)*dil) still lies within the padded image.
  int lower[2] = {-pad_w, -pad_h};
  int upper[2] = {pad_w - (S - 1) * dil_w, pad_h - (R - 1) * dil_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1};
  CUresult r = g_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                        upper, channels, pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DL_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d): NHWC=%d,%d,%d,%d ldx=%d RS=%dx%d", (int)r, N, H,
                W, C, ldx, R, S);
  // Drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (a flag in the
  // second descriptor word that must be clear); same remedy as NVIDIA's own conv templates apply.
  if (g_driver_version <= 13010) {
    uint64_t bytes = (uint64_t)N * img_rows * img_cols * ldx * 2;
    if (bytes < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  }
  return DL_OK;
}

}  // namespace dl

namespace dl {
static std::atomic<int> g_opt_pair{1}, g_opt_pair_resident{1}, g_opt_dbg{0}, g_opt_tap_share{1}, g_opt_frontend{2},
    g_opt_stft_pad{0}, g_opt_staged{1}, g_opt_prepass{2}, g_opt_small_linear{1}, g_opt_statpool_mlp{4}, g_opt_statpool_slab{256}, g_opt_stem{2}, g_opt_pool_fuse{1};
int opt_pair() { return g_opt_pair.load(std::memory_order_relaxed); }
int opt_pair_resident() { return g_opt_pair_resident.load(std::memory_order_relaxed); }
int opt_staged_epilogue() { return g_opt_staged.load(std::memory_order_relaxed); }
int opt_tap_share() { return g_opt_tap_share.load(std::memory_order_relaxed); }
int opt_dbg() { return g_opt_dbg.load(std::memory_order_relaxed); }
int opt_frontend() { return g_opt_frontend.load(std::memory_order_relaxed); }
int opt_prepass() { return g_opt_prepass.load(std::memory_order_relaxed); }
int opt_small_linear() { return g_opt_small_linear.load(std::memory_order_relaxed); }
int opt_statpool_mlp() { return g_opt_statpool_mlp.load(std::memory_order_relaxed); }
int opt_statpool_slab() { return g_opt_statpool_slab.load(std::memory_order_relaxed); }
int opt_stft_pad() { return g_opt_stft_pad.load(std::memory_order_relaxed); }
int opt_stem() { return g_opt_stem.load(std::memory_order_relaxed); }
int opt_pool_fuse() { return g_opt_pool_fuse.load(std::memory_order_relaxed); }
}  // namespace dl

extern "C" {
int dl_set_option(const char* name, int value) {
  if (!name) return DL_ERR_INVALID;
  if (!strcmp(name, "pair")) { dl::g_opt_pair.store(value); return DL_OK; }
  if (!strcmp(name, "tap_share")) { dl::g_opt_tap_share.store(value); return DL_OK; }
  if (!strcmp(name, "dbg")) { dl::g_opt_dbg.store(value); return DL_OK; }
  if (!strcmp(name, "frontend")) { dl::g_opt_frontend.store(value); return DL_OK; }
  if (!strcmp(name, "prepass")) { dl::g_opt_prepass.store(value); return DL_OK; }
  if (!strcmp(name, "small_linear")) { dl::g_opt_small_linear.store(value); return DL_OK; }
  if (!strcmp(name, "statpool_mlp")) { dl::g_opt_statpool_mlp.store(value); return DL_OK; }
  if (!strcmp(name, "statpool_slab")) { dl::g_opt_statpool_slab.store(value); return DL_OK; }
  if (!strcmp(name, "pool_fuse")) { dl::g_opt_pool_fuse.store(value); return DL_OK; }
  if (!strcmp(name, "stem")) { dl::g_opt_stem.store(value); return DL_OK; }
  if (!strcmp(name, "stft_pad")) { dl::g_opt_stft_pad.store(value); return DL_OK; }
  if (!strcmp(name, "pair_resident")) { dl::g_opt_pair_resident.store(value); return DL_OK; }
  if (!strcmp(name, "staged_epilogue")) { dl::g_opt_staged.store(value); return DL_OK; }
  return dl::fail(DL_ERR_INVALID, "unknown option '%s'", name);
}
int dl_version(void) { return 100; }
const char* dl_last_error(void) { return dl::g_err; }
long long dl_launch_count(void) { return dl::g_launches.load(); }
}
