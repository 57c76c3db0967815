// Runs the SOURCE of the generation-2 front-end kernels (deeplip_b200/csrc/frontend_gen2.cuh) on the CPU: one OS thread
// per CUDA thread, std::barrier for __syncthreads / __syncwarp, an exchange buffer for the warp shuffles.  Test
// infrastructure only (tests/test_host_logic.py::test_frontend_kernel_source_on_cpu_threads compares the output with
// oracle/frontend_np.py); it checks index arithmetic, framing, tables and staging -- not memory-model behaviour.
//   usage: frontend_cpu_emul kind F nsamp B pad_mode cmvn+10*delta wav.f32 lengths.i32|- out.bin
//   out.bin = feat_f32 (B,Fout,T) float32 followed by feat_bf16 (B,T,ld) uint16, Fout = F (1 + delta), ld = ceil64(Fout)
//   usage: frontend_cpu_emul prepass is_u8 frames H W Hraw Wraw in.bin out.bin     (stem_prepass.cuh; mean .421 std .165)
//   out.bin = (frames, H+8, pitch) uint16 bf16, pitch = ceil8(W+8)
//   usage: frontend_cpu_emul linear M C Cout ldx ldw with_bf16 with_scale2 in.bin out.bin     (linear_small.cuh)
//   in.bin = x (M,ldx) u16 | w (Cout,ldw) u16 | scale, shift, slope, scale2, shift2 (Cout f32 each); f32_slope = 0.2
//   out.bin = y (M,Cout) u16 | yf (M,Cout) f32
//   usage: frontend_cpu_emul pool B T HW C in.bin out.bin                                    (pool_kernels.cuh)
//   in.bin = lengths (B i32) | x (B*T, HW, C) u16 bf16
//   out.bin = frame_feats (B,T,C) f32 | utt_mean (B,C) f32 [frame_pool_kernel] | utt_mean (B,C) f32 [temporal_mean_kernel
//             on those frame features]
//   usage: frontend_cpu_emul plda n_utt D R n_trials c0 in.bin out.bin - -                   (plda_score.cuh)
//   in.bin = emb (n_utt,D) f32 | M (R,D) f32 | bias, k1, k2 (R f32 each) | enrol, test (n_trials i32 each)
//   out.bin = u (n_utt,R) f32 | scores (n_trials) f32
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__

struct Idx3 { int x, y, z; };
thread_local Idx3 threadIdx, blockIdx;
Idx3 gridDim, blockDim;
struct uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#define DL_STATIC_SHARED static          // a kernel's static shared array: one per process here (blocks run one after another)
using std::max;
using std::min;
template <typename T> inline T __ldg(const T* p) { return *p; }
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31));
}

struct BlockCtx {
  std::barrier<> all;
  std::vector<std::unique_ptr<std::barrier<>>> warp;
  float xch[16][32];
  explicit BlockCtx(int nthreads) : all(nthreads) { for (int i = 0; i < 16; ++i) warp.emplace_back(new std::barrier<>(32)); }
};
thread_local BlockCtx* g_ctx;
inline void __syncthreads() { g_ctx->all.arrive_and_wait(); }
inline void __syncwarp() { g_ctx->warp[threadIdx.x >> 5]->arrive_and_wait(); }

namespace dl {
inline float warp_sum(float v) {       // the xor butterfly of dl_ptx.cuh
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int off = 16; off > 0; off >>= 1) {
    g_ctx->xch[w][lane] = v;
    __syncwarp();
    const float o = g_ctx->xch[w][lane ^ off];
    __syncwarp();
    v += o;
  }
  return v;
}
inline uint32_t bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return u >> 16;
}
inline uint32_t pack_bf16x2(float lo, float hi) { return bf16_rn(lo) | (bf16_rn(hi) << 16); }
inline float bf16_lo(uint32_t v) { v <<= 16; float f; memcpy(&f, &v, 4); return f; }
inline float bf16_hi(uint32_t v) { v &= 0xffff0000u; float f; memcpy(&f, &v, 4); return f; }
alignas(16) float smf[64 * 1024];      // the kernels' dynamic shared memory
alignas(16) float rows[64 * 1024];
alignas(16) float lin_part[16 * 4 * 64];
}  // namespace dl

#include "../deeplip_b200/csrc/frontend_gen2.cuh"
#include "../deeplip_b200/csrc/stem_prepass.cuh"
#include "../deeplip_b200/csrc/linear_small.cuh"
#include "../deeplip_b200/csrc/plda_score.cuh"
#include "../deeplip_b200/csrc/pool_kernels.cuh"

template <typename Fn>
static void launch(int gx, int gy, Fn fn, int nthreads = 256) {
  gridDim = Idx3{gx, gy, 1};
  blockDim = Idx3{nthreads, 1, 1};
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx) {
      BlockCtx ctx(nthreads);
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
          threadIdx = Idx3{t, 0, 0};
          blockIdx = Idx3{bx, by, 0};
          g_ctx = &ctx;
          fn();
        });
      for (auto& x : th) x.join();
    }
}

static int prepass_main(char** a) {
  const int is_u8 = atoi(a[2]), frames = atoi(a[3]), H = atoi(a[4]), W = atoi(a[5]), Hraw = atoi(a[6]), Wraw = atoi(a[7]);
  const int rows = H + 8, pitch = (W + 8 + 7) / 8 * 8;
  const size_t nin = (size_t)frames * Hraw * Wraw * (is_u8 ? 1 : 4);
  std::vector<uint8_t> in(nin + 4);
  FILE* f = fopen(a[8], "rb");
  if (!f || fread(in.data(), 1, nin, f) != nin) return 3;
  fclose(f);
  std::vector<uint16_t> out((size_t)frames * rows * pitch, 0x7fc0);
  const float mean = 0.421f, std_ = 0.165f;
  const int dh = is_u8 ? (Hraw - H) / 2 : 0, dw = is_u8 ? (Wraw - W) / 2 : 0;
  const void* x = in.data();
  uint16_t* xp = out.data();
  const int aligned4 = (Wraw % 4 == 0 && ((uintptr_t)x & 3) == 0) ? 1 : 0;
  // frames = clips x 2; with 4 frames the second clip is ragged (length 1): its frame 1 must come out as zeros
  const int Tclip = 2;
  std::vector<int32_t> lv(frames / 2 + 1, 2);
  const int32_t* lens = nullptr;
  if (frames == 4) { lv[1] = 1; lens = lv.data(); }
  launch(frames, 1, [&] {
    dl::stem_prepass2_kernel(x, is_u8, H, W, Hraw, Wraw, dh, dw, is_u8 ? 1.0f / (255.0f * std_) : 1.0f,
                             is_u8 ? -mean / std_ : 0.0f, rows, pitch, aligned4, Tclip, lens, xp);
  });
  f = fopen(a[9], "wb");
  fwrite(out.data(), 2, out.size(), f);
  fclose(f);
  printf("rows %d pitch %d aligned %d\n", rows, pitch, aligned4);
  return 0;
}

static int linear_main(char** a) {
  using namespace dl;
  const int M = atoi(a[2]), C = atoi(a[3]), Cout = atoi(a[4]), ldx = atoi(a[5]), ldw = atoi(a[6]), wb = atoi(a[7]),
            ws2 = atoi(a[8]);
  std::vector<uint16_t> x((size_t)M * ldx), w((size_t)Cout * ldw), y((size_t)M * Cout, 0x7fc0);
  std::vector<float> prm((size_t)5 * Cout), yf((size_t)M * Cout, -777.f);
  FILE* f = fopen(a[9], "rb");
  if (!f || fread(x.data(), 2, x.size(), f) != x.size() || fread(w.data(), 2, w.size(), f) != w.size() ||
      fread(prm.data(), 4, prm.size(), f) != prm.size()) return 3;
  fclose(f);
  LinearSmallParams p{};
  p.x = x.data(); p.w = w.data(); p.M = M; p.C = C; p.ldx = ldx; p.ldw = ldw; p.Cout = Cout;
  p.scale = prm.data(); p.shift = prm.data() + Cout; p.slope = prm.data() + 2 * Cout;
  p.y = wb ? y.data() : nullptr; p.ldy = Cout;
  p.scale2 = ws2 ? prm.data() + 3 * Cout : nullptr; p.shift2 = ws2 ? prm.data() + 4 * Cout : nullptr;
  p.f32_slope = 0.2f; p.yf = yf.data(); p.ldf = Cout;
  const int grid = (Cout + kLinCh - 1) / kLinCh;
  if (M <= 32) launch(grid, 1, [&] { linear_small_kernel<1>(p); }, 32 * kLinKs);
  else launch(grid, (M + 63) / 64, [&] { linear_small_kernel<2>(p); }, 32 * kLinKs);
  f = fopen(a[10], "wb");
  fwrite(y.data(), 2, y.size(), f);
  fwrite(yf.data(), 4, yf.size(), f);
  fclose(f);
  return 0;
}

static int pool_main(char** a) {
  using namespace dl;
  const int B = atoi(a[2]), T = atoi(a[3]), HW = atoi(a[4]), C = atoi(a[5]);
  std::vector<int32_t> len(B);
  std::vector<uint16_t> x((size_t)B * T * HW * C);
  FILE* f = fopen(a[6], "rb");
  if (!f || fread(len.data(), 4, len.size(), f) != len.size() || fread(x.data(), 2, x.size(), f) != x.size()) return 3;
  fclose(f);
  std::vector<float> ff((size_t)B * T * C, -777.f), um((size_t)B * C, -777.f), um2((size_t)B * C, -777.f);
  const uint16_t* px = x.data();
  const int32_t* pl = len.data();
  float *pf = ff.data(), *pu = um.data(), *pu2 = um2.data();
  if (HW == 9) launch(B, (C + 63) / 64, [&] { frame_pool_kernel<9>(px, T, HW, C, pl, pf, pu); });
  else launch(B, (C + 63) / 64, [&] { frame_pool_kernel<0>(px, T, HW, C, pl, pf, pu); });
  launch(B, (C + 63) / 64, [&] { temporal_mean_kernel(pf, T, C, pl, pu2); });
  f = fopen(a[7], "wb");
  fwrite(ff.data(), 4, ff.size(), f);
  fwrite(um.data(), 4, um.size(), f);
  fwrite(um2.data(), 4, um2.size(), f);
  fclose(f);
  return 0;
}

static int plda_main(char** a) {
  using namespace dl;
  const int n_utt = atoi(a[2]), D = atoi(a[3]), R = atoi(a[4]), nt = atoi(a[5]);
  const float c0 = (float)atof(a[6]);
  std::vector<float> emb((size_t)n_utt * D), M((size_t)R * D), prm((size_t)3 * R), u((size_t)n_utt * R, -777.f), sc(nt, -777.f);
  std::vector<int32_t> idx((size_t)2 * nt);
  FILE* f = fopen(a[7], "rb");
  if (!f || fread(emb.data(), 4, emb.size(), f) != emb.size() || fread(M.data(), 4, M.size(), f) != M.size() ||
      fread(prm.data(), 4, prm.size(), f) != prm.size() || fread(idx.data(), 4, idx.size(), f) != idx.size()) return 3;
  fclose(f);
  const float *pe = emb.data(), *pm = M.data(), *pb = prm.data(), *k1 = prm.data() + R, *k2 = prm.data() + 2 * R;
  float *pu = u.data(), *ps = sc.data();
  const int32_t *en = idx.data(), *te = idx.data() + nt;
  launch((n_utt + 7) / 8, 1, [&] { plda_transform_kernel(pe, n_utt, D, pm, pb, R, pu); });
  launch((nt + 255) / 256, 1, [&] { plda_llr_trials_kernel(pu, n_utt, R, k1, k2, c0, en, te, nt, ps); });
  f = fopen(a[8], "wb");
  fwrite(u.data(), 4, u.size(), f);
  fwrite(sc.data(), 4, sc.size(), f);
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  using namespace dl;
  if (argc == 8 && !strcmp(argv[1], "pool")) return pool_main(argv);
  if (argc == 11 && !strcmp(argv[1], "plda")) return plda_main(argv);
  if (argc == 11 && !strcmp(argv[1], "linear")) return linear_main(argv);
  if (argc != 10) return 2;
  if (!strcmp(argv[1], "prepass")) return prepass_main(argv);
  const int kind = atoi(argv[1]), F = atoi(argv[2]), nsamp = atoi(argv[3]), B = atoi(argv[4]), pad = atoi(argv[5]),
            cmvn = atoi(argv[6]) % 10, delta = atoi(argv[6]) / 10, Fout = F * (1 + delta);
  const bool stft = kind == 3;
  const int T = stft ? 1 + nsamp / kFrameStep
                     : (nsamp <= kFrameLen ? 1 : 1 + (nsamp - kFrameLen + kFrameStep - 1) / kFrameStep);
  const int ld = (Fout + 63) / 64 * 64;
  std::vector<float> wav((size_t)B * nsamp), feat((size_t)B * Fout * T, -777.f);
  std::vector<uint16_t> b16((size_t)B * T * ld, 0x7fc0);
  std::vector<int32_t> len(B);
  FILE* f = fopen(argv[7], "rb");
  if (!f || fread(wav.data(), 4, wav.size(), f) != wav.size()) return 3;
  fclose(f);
  const int32_t* lengths = nullptr;
  if (strcmp(argv[8], "-")) {
    f = fopen(argv[8], "rb");
    if (!f || fread(len.data(), 4, B, f) != (size_t)B) return 3;
    fclose(f);
    lengths = len.data();
  }
  if ((size_t)frames2_smem_floats(F) > sizeof(smf) / 4 || (size_t)8 * T > sizeof(rows) / 4) return 4;
  fill_frontend_const(&g_fc);
  FrontendTables tb;
  fill_frontend_tables(&tb, kind, F, pad);
  const float* w = wav.data();
  float* ft = feat.data();
  uint16_t* ob = b16.data();
  if (stft) launch((T + kBlkFrames - 1) / kBlkFrames, B, [&] { frontend_frames2_kernel<true>(w, lengths, nsamp, T, tb, ft, Fout); });
  else launch((T + kBlkFrames - 1) / kBlkFrames, B, [&] { frontend_frames2_kernel<false>(w, lengths, nsamp, T, tb, ft, Fout); });
  launch(B, (F + 7) / 8, [&] { frontend_cmvn2_kernel(ft, lengths, nsamp, T, F, cmvn, stft ? 1 : 0, delta, ob, ld); });
  f = fopen(argv[9], "wb");
  fwrite(feat.data(), 4, feat.size(), f);
  fwrite(b16.data(), 2, b16.size(), f);
  fclose(f);
  printf("T %d ld %d\n", T, ld);
  return 0;
}
