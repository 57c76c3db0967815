// K1: fused audio front end -- pre-emphasis + 400/160 framing + 512-point FFT power spectrum + triangular
// mel filterbank + log (+ DCT-II, lifter, log-energy for MFCC), or the `stft` feature (centre-padded Hann frames,
// log1p |S|), then per-utterance CMVN.
// Restates python_speech_features v0.6 / librosa.stft as the reference calls them
// (models/fusion_models/datasets.py:227-246) and `_normalize` (:214-215).  4 B/sample in, F*T*(4+2) B out.
// Two generations live here: frontend_frames2/cmvn2 (default: register-resident radix-8 FFT, fft512.cuh) and the
// first radix-2 shared-memory kernels (dl_set_option("frontend", 1)) kept for A/B measurements.
#include <math.h>
#include <string.h>
#include "dl_host.cuh"
#include "dl_ptx.cuh"
#include "frontend_gen2.cuh"

namespace dl {

__constant__ float2 c_twiddle[kNfft / 2];   // exp(-2 pi i k / 512)

__device__ __forceinline__ int bitrev9(int v) { return __brev((unsigned)v) >> 23; }

// grid (ceil(T/16), B); block 256 = 8 warps, two frames per warp (two-for-one real FFT).  Writes raw (pre-CMVN) features to
// feat (B, F, T) f32.
__global__ void __launch_bounds__(256, 5) frontend_frames_kernel(const float* __restrict__ wav,
                                                              const int32_t* __restrict__ lengths, int nsamp, int T,
                                                              FrontendTables tb, float* __restrict__ feat) {
  __shared__ float2 buf[8][kNfft];
  __shared__ float2 tw[kNfft / 2];        // twiddles staged in shared memory: lanes index them divergently, which the
                                          // constant cache would serialise 32-fold
  __shared__ float logmel[8][kMaxFilt];
  __shared__ float logmel2[8][kMaxFilt];
  __shared__ float dctm[26 * 26];         // DCT rows (ncep <= 26 x 26 filters), mfcc only; 40.6 KB in all -> 5 CTAs / SM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t0 = (blockIdx.x * 8 + warp) * 2;       // this warp transforms frames t0 and t0+1 with ONE complex FFT
  const int F = tb.ncep;

  for (int i = threadIdx.x; i < kNfft / 2; i += 256) tw[i] = c_twiddle[i];
  if (tb.kind == 0) {
    // scipy dct(type=2, norm='ortho') rows with the cepstral lifter (L=22) folded in
    const int nf = tb.nfilt;
    for (int i = threadIdx.x; i < F * nf; i += 256) {
      const int n = i / nf, j = i % nf;
      const float ortho = (n == 0) ? sqrtf(1.f / (float)nf) : sqrtf(2.f / (float)nf);
      const float lift = 1.f + 11.f * sinpif((float)n / 22.f);
      dctm[i] = ortho * lift * cospif((float)(n * (2 * j + 1)) / (float)(2 * nf));
    }
  }
  __syncthreads();

  const int len = lengths ? min(lengths[b], nsamp) : nsamp;
  const int nfr = len <= kFrameLen ? 1 : 1 + (len - kFrameLen + kFrameStep - 1) / kFrameStep;
  if (t0 >= T) return;
  const bool has2 = t0 + 1 < T;
  float* out = feat + (size_t)b * F * T + t0;
  if (t0 >= nfr) {   // padding frames of a ragged batch
    for (int f = lane; f < F; f += 32) {
      out[(size_t)f * T] = 0.f;
      if (has2) out[(size_t)f * T + 1] = 0.f;
    }
    return;
  }

  // ---- load + pre-emphasis + zero pad to 512, bit-reversed for the in-place DIT FFT.
  // Two real frames ride one complex transform: z = x1 + i x2  =>  X1[k] = (Z[k] + conj Z[N-k]) / 2,
  // X2[k] = (Z[k] - conj Z[N-k]) / (2i).
  const float* x = wav + (size_t)b * nsamp;
  const int s0 = t0 * kFrameStep;
  const bool live2 = t0 + 1 < nfr;
  float2* z = buf[warp];
  for (int i = lane; i < kNfft; i += 32) {
    float v1 = 0.f, v2 = 0.f;
    if (i < kFrameLen) {
      const int s = s0 + i;
      if (s < len) {
        const float cur = __ldg(x + s);
        v1 = (s == 0) ? cur : cur - kPreemph * __ldg(x + s - 1);
      }
      const int s2 = s + kFrameStep;
      if (live2 && s2 < len) v2 = __ldg(x + s2) - kPreemph * __ldg(x + s2 - 1);
    }
    z[bitrev9(i)] = make_float2(v1, v2);
  }
  __syncwarp();
  // ---- 9 radix-2 stages
#pragma unroll 1
  for (int st = 0; st < 9; ++st) {
    const int half = 1 << st;
    for (int i = lane; i < kNfft / 2; i += 32) {
      const int grp = i >> st, pos = i & (half - 1);
      const int i0 = (grp << (st + 1)) + pos, i1 = i0 + half;
      const float2 w = tw[pos << (8 - st)];
      const float2 a = z[i0], c = z[i1];
      const float2 m = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
      z[i0] = make_float2(a.x + m.x, a.y + m.y);
      z[i1] = make_float2(a.x - m.x, a.y - m.y);
    }
    __syncwarp();
  }
  // ---- split the two spectra; power (257 bins) of frame 1 / frame 2 into z[k].x / z[k].y; total energies
  float e1 = 0.f, e2 = 0.f;
  const float inv_n = 1.f / (float)kNfft;
  for (int k = lane + 1; k < kNfft / 2; k += 32) {
    const float2 a = z[k], c = z[kNfft - k];
    const float x1r = 0.5f * (a.x + c.x), x1i = 0.5f * (a.y - c.y);
    const float x2r = 0.5f * (a.y + c.y), x2i = 0.5f * (c.x - a.x);
    const float p1 = (x1r * x1r + x1i * x1i) * inv_n, p2 = (x2r * x2r + x2i * x2i) * inv_n;
    e1 += p1; e2 += p2;
    z[k] = make_float2(p1, p2);
  }
  if (lane == 0) {
    const float2 a = z[0], c = z[kNfft / 2];
    const float p10 = a.x * a.x * inv_n, p20 = a.y * a.y * inv_n;
    const float p1n = c.x * c.x * inv_n, p2n = c.y * c.y * inv_n;
    e1 += p10 + p1n; e2 += p20 + p2n;
    z[0] = make_float2(p10, p20);
    z[kNfft / 2] = make_float2(p1n, p2n);
  }
  e1 = warp_sum(e1); e2 = warp_sum(e2);
  if (e1 == 0.f) e1 = 2.220446049250313e-16f;
  if (e2 == 0.f) e2 = 2.220446049250313e-16f;
  __syncwarp();
  // ---- mel filterbank: lane j <-> filter j (two passes when nfilt > 32), both frames at once
  float* lm1 = logmel[warp];
  float* lm2 = logmel2[warp];
  for (int j = lane; j < tb.nfilt; j += 32) {
    const int b0 = tb.bins[j], b1 = tb.bins[j + 1], b2 = tb.bins[j + 2];
    float s1 = 0.f, s2 = 0.f;
    const float up = tb.inv_width[j], dn = tb.inv_width[j + 1];
    for (int i = b0; i < b1; ++i) {
      const float w = (float)(i - b0) * up;
      const float2 pw = z[i];
      s1 = fmaf(pw.x, w, s1); s2 = fmaf(pw.y, w, s2);
    }
    for (int i = b1; i < b2; ++i) {
      const float w = (float)(b2 - i) * dn;
      const float2 pw = z[i];
      s1 = fmaf(pw.x, w, s1); s2 = fmaf(pw.y, w, s2);
    }
    if (s1 == 0.f) s1 = 2.220446049250313e-16f;
    if (s2 == 0.f) s2 = 2.220446049250313e-16f;
    lm1[j] = (tb.kind == 1) ? s1 : logf(s1);
    lm2[j] = (tb.kind == 1) ? s2 : logf(s2);
  }
  __syncwarp();
  if (tb.kind == 0) {
    for (int n = lane; n < F; n += 32) {
      float c1 = 0.f, c2 = 0.f;
      for (int j = 0; j < tb.nfilt; ++j) {
        const float d = dctm[n * tb.nfilt + j];
        c1 = fmaf(d, lm1[j], c1); c2 = fmaf(d, lm2[j], c2);
      }
      if (n == 0) { c1 = logf(e1); c2 = logf(e2); }          // appendEnergy=True
      out[(size_t)n * T] = c1;
      if (has2) out[(size_t)n * T + 1] = live2 ? c2 : 0.f;
    }
  } else {
    for (int f = lane; f < F; f += 32) {
      out[(size_t)f * T] = lm1[f];
      if (has2) out[(size_t)f * T + 1] = live2 ? lm2[f] : 0.f;
    }
  }
}

// grid (B, ceil(F/8)); block 256: one coefficient row per warp (biased std, + 2e-12), writes f32 in place
// and the channels-last bf16 copy the TDNN consumes.
__global__ void __launch_bounds__(256) frontend_cmvn_kernel(float* __restrict__ feat, const int32_t* __restrict__ lengths,
                                                            int nsamp, int T, int F, int cmvn,
                                                            uint16_t* __restrict__ out_bf16, int ld) {
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lengths ? min(lengths[b], nsamp) : nsamp;
  int nfr = len <= kFrameLen ? 1 : 1 + (len - kFrameLen + kFrameStep - 1) / kFrameStep;
  nfr = min(nfr, T);
  for (int f = blockIdx.y * 8 + warp; f < F; f += 8 * gridDim.y) {
    float* row = feat + ((size_t)b * F + f) * T;
    float mean = 0.f, inv = 1.f;
    if (cmvn) {
      float s = 0.f;
      for (int t = lane; t < nfr; t += 32) s += row[t];
      mean = warp_sum(s) / (float)nfr;
      float q = 0.f;
      for (int t = lane; t < nfr; t += 32) { const float d = row[t] - mean; q = fmaf(d, d, q); }
      q = warp_sum(q) / (float)nfr;
      inv = 1.f / (sqrtf(q) + 2e-12f);
    }
    for (int t = lane; t < T; t += 32) {
      const float v = t < nfr ? (row[t] - mean) * inv : 0.f;
      row[t] = v;
      if (out_bf16) reinterpret_cast<__nv_bfloat16*>(out_bf16)[((size_t)b * T + t) * ld + f] = __float2bfloat16_rn(v);
    }
  }
  if (out_bf16) {   // zero the padded channels
    for (int i = blockIdx.y * 256 + threadIdx.x; i < T * (ld - F); i += 256 * gridDim.y) {
      const int t = i / (ld - F), c = F + i % (ld - F);
      reinterpret_cast<__nv_bfloat16*>(out_bf16)[((size_t)b * T + t) * ld + c] = __float2bfloat16_rn(0.f);
    }
  }
}


static int upload_twiddles() {
  static PerDevice<bool> done_dev;
  bool* done = done_dev.slot();
  if (!done) return fail(DL_ERR_CUDA, "frontend: no current device");
  if (*done) return DL_OK;
  float2 tw[kNfft / 2];
  for (int k = 0; k < kNfft / 2; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)kNfft;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  cudaError_t e = cudaMemcpyToSymbol(c_twiddle, tw, sizeof(tw));
  if (e != cudaSuccess) return fail(DL_ERR_CUDA, "frontend twiddles: %s", cudaGetErrorString(e));
  *done = true;
  return DL_OK;
}

// FFT twiddles, DCT matrix and Hann window of generation 2, once per device.
static int upload_tables() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(DL_ERR_CUDA, "frontend: cudaGetDevice failed");
  if (done[dev]) return DL_OK;
  static FrontendConst h;
  fill_frontend_const(&h);
  cudaError_t e = cudaMemcpyToSymbol(g_fc, &h, sizeof(h));
  if (e != cudaSuccess) return fail(DL_ERR_CUDA, "frontend tables: %s", cudaGetErrorString(e));
  e = cudaFuncSetAttribute(frontend_frames2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           frames2_smem_floats(kMaxFilt) * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(frontend_frames2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             frames2_smem_floats(kNfft / 2 + 1) * 4);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(frontend_cmvn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCmvn2MaxSmem);
  if (e != cudaSuccess) return fail(DL_ERR_CUDA, "frontend smem attribute: %s", cudaGetErrorString(e));
  done[dev] = true;
  return DL_OK;
}

}  // namespace dl

static int frontend_features_impl(const float* wav, int pcm16, const int32_t* lengths, int B, int nsamp, int kind, int F,
                                  int cmvn, int delta, void* feat_bf16, int ld_bf16, float* feat_f32, int T,
                                  void* stream);

extern "C" int dl_frontend_features(const float* wav, const int32_t* lengths, int B, int nsamp, int kind, int F,
                                    int cmvn, int delta, void* feat_bf16, int ld_bf16, float* feat_f32, int T,
                                    void* stream) {
  return frontend_features_impl(wav, 0, lengths, B, nsamp, kind, F, cmvn, delta, feat_bf16, ld_bf16, feat_f32, T, stream);
}

extern "C" int dl_frontend_features_pcm16(const int16_t* wav, const int32_t* lengths, int B, int nsamp, int kind, int F,
                                          int cmvn, int delta, void* feat_bf16, int ld_bf16, float* feat_f32, int T,
                                          void* stream) {
  return frontend_features_impl(reinterpret_cast<const float*>(wav), 1, lengths, B, nsamp, kind, F, cmvn, delta, feat_bf16,
                                ld_bf16, feat_f32, T, stream);
}

static int frontend_features_impl(const float* wav, int pcm16, const int32_t* lengths, int B, int nsamp, int kind, int F,
                                  int cmvn, int delta, void* feat_bf16, int ld_bf16, float* feat_f32, int T,
                                  void* stream) {
  using namespace dl;
  DL_CHECK_ARG(wav && feat_f32, "frontend: wav and feat_f32 are required");
  DL_CHECK_ARG(B > 0 && nsamp > 0, "frontend: empty batch");
  DL_CHECK_ARG(kind >= 0 && kind <= 3, "frontend: kind must be 0 (mfcc), 1 (fbank), 2 (logfbank) or 3 (stft)");
  const bool stft = kind == 3;
  const int nfilt = kind == 0 ? 26 : F;
  if (stft) {
    DL_CHECK_ARG(F == kNfft / 2 + 1, "frontend: stft has F = 257 bins (n_fft = 512), got %d", F);
  } else {
    DL_CHECK_ARG(F >= 1 && F <= kMaxFilt && nfilt <= kMaxFilt && F <= nfilt, "frontend: F=%d out of range", F);
  }
  const int Texp = stft ? 1 + nsamp / kFrameStep
                        : (nsamp <= kFrameLen ? 1 : 1 + (nsamp - kFrameLen + kFrameStep - 1) / kFrameStep);
  DL_CHECK_ARG(T == Texp, "frontend: T=%d but %d samples give %d frames", T, nsamp, Texp);
  DL_CHECK_ARG(delta >= 0 && delta <= 2, "frontend: delta order must be 0, 1 or 2");
  const int Fout = F * (1 + delta);
  DL_CHECK_ARG(!feat_bf16 || ld_bf16 >= Fout, "frontend: ld_bf16 < F (1 + delta)");
  const int gen = opt_frontend();
  DL_CHECK_ARG(delta == 0 || gen >= 2, "frontend: delta features need the generation-2 kernels");
  DL_CHECK_ARG(!stft || gen >= 2, "frontend: stft needs the generation-2 kernels (dl_set_option(\"frontend\", 2))");
  DL_CHECK_ARG(!pcm16 || gen >= 2, "frontend: int16 PCM input needs the generation-2 kernels");

  FrontendTables tb;
  fill_frontend_tables(&tb, kind, F, opt_stft_pad());
  tb.pcm16 = pcm16;
  cudaStream_t s = (cudaStream_t)stream;
  int st;
  if (gen >= 2) {
    st = upload_tables();
    if (st != DL_OK) return st;
    dim3 grid((T + kBlkFrames - 1) / kBlkFrames, B);
    const size_t smem = (size_t)frames2_smem_floats(F) * 4;
    if (stft) frontend_frames2_kernel<true><<<grid, 256, smem, s>>>(wav, lengths, nsamp, T, tb, feat_f32, Fout);
    else frontend_frames2_kernel<false><<<grid, 256, smem, s>>>(wav, lengths, nsamp, T, tb, feat_f32, Fout);
    st = check_launch("frontend_frames2_kernel");
  } else {
    st = upload_twiddles();
    if (st != DL_OK) return st;
    dim3 grid((T + 15) / 16, B);
    frontend_frames_kernel<<<grid, 256, 0, s>>>(wav, lengths, nsamp, T, tb, feat_f32);
    st = check_launch("frontend_frames_kernel");
  }
  if (st != DL_OK) return st;
  const size_t csmem = (size_t)8 * T * 4;
  const bool bf16_ok = !feat_bf16 || (ld_bf16 % 8 == 0 && ld_bf16 >= (Fout + 7) / 8 * 8 && ((uintptr_t)feat_bf16 & 15) == 0);
  if (gen >= 2 && csmem <= (size_t)kCmvn2MaxSmem && bf16_ok) {
    frontend_cmvn2_kernel<<<dim3(B, (F + 7) / 8), 256, csmem, s>>>(feat_f32, lengths, nsamp, T, F, cmvn, stft ? 1 : 0, delta,
                                                                  (uint16_t*)feat_bf16, ld_bf16);
    return check_launch("frontend_cmvn2_kernel");
  }
  DL_CHECK_ARG(delta == 0, "frontend: delta features need ld_bf16 %% 8 == 0 >= ceil8(F (1 + delta)), a 16-byte aligned feat_bf16 and T <= %d", kCmvn2MaxSmem / 32);
  DL_CHECK_ARG(!stft, "frontend: stft needs ld_bf16 %% 8 == 0, a 16-byte aligned feat_bf16 and T <= %d", kCmvn2MaxSmem / 32);
  frontend_cmvn_kernel<<<dim3(B, (F + 7) / 8), 256, 0, s>>>(feat_f32, lengths, nsamp, T, F, cmvn, (uint16_t*)feat_bf16, ld_bf16);
  return check_launch("frontend_cmvn_kernel");
}
