// Generation-2 audio front end kernels (K1) and the host code that fills their tables.  Kept in a header with no CUDA
// runtime dependency beyond the usual built-ins so that tests/frontend_cpu_emul.cpp can compile THIS SOURCE for the CPU
// (one OS thread per CUDA thread, barriers for __syncthreads / __syncwarp) and check it against the oracle without a GPU.
// The includer provides: warp_sum(float), pack_bf16x2(float, float), __ldg, uint4 / make_uint4, min / max.
#pragma once
#include <math.h>
#include <stdint.h>
#include "fft512.cuh"

namespace dl {

constexpr int kNfft = 512;
constexpr int kFrameLen = 400;
constexpr int kFrameStep = 160;
constexpr int kMaxFilt = 64;
constexpr float kPreemph = 0.97f;

struct FrontendTables {
  int nfilt;
  int ncep;                  // number of coefficients written (F)
  int kind;                  // 0 mfcc, 1 fbank, 2 logfbank, 3 stft
  int pad_mode;              // stft: 0 reflect (librosa < 0.10 default), 1 zeros (librosa >= 0.10 default)
  int pcm16;                 // 1: wav is int16 PCM, sample = value / 32768 (what soundfile.read returns for a 16-bit file)
  int bins[kMaxFilt + 2];    // FFT-bin edges of the triangular filters
  float inv_width[kMaxFilt + 1];   // 1 / (bins[j+1] - bins[j])
};

// Per-device constant tables, written once (upload_tables): FFT twiddles, the compact pass-2 twiddles, the 26 x 26
// ortho DCT-II matrix with the cepstral lifter folded in, the periodic Hann window of 400 centred in 512.
struct FrontendConst {
  Cx<float> tw[kNfft];
  Cx<float> tw64[kFftTw64];
  float dct[26 * 26];
  float hann[kNfft];
};
__device__ FrontendConst g_fc;

constexpr int kBlkFrames = 16;                                        // frames per block: 8 warps x 2
constexpr int kChunk = (kBlkFrames - 1) * kFrameStep + kNfft;         // samples a block stages (2912)
constexpr int kOPitch = kBlkFrames + 1;
constexpr int kCmvn2MaxSmem = 160 * 1024;                             // 8 rows x T <= 5120 frames                               // output staging pitch (bank-conflict free)

__host__ __device__ constexpr int frames2_smem_floats(int F) {
  return 2 * 8 * kFftScratch + 2 * kNfft + 2 * kFftTw64 + kChunk + 704 + 68 + 68 + 8 * 2 * kMaxFilt + F * kOPitch;
}

// grid (ceil(T/16), B); block 256 = 8 warps.  The block stages its 2912 (pre-emphasised / reflect-padded) samples in
// shared memory once; each warp transforms frames t0, t0+1 as ONE complex FFT held in registers (16 points per lane,
// three radix-8 passes); results of the 16 frames are staged and leave as 64-byte row segments of rows [0, F) of
// feat (B, feat_rows, T) f32.
template <bool kStft>
__global__ void __launch_bounds__(256, 2) frontend_frames2_kernel(const float* __restrict__ wav,
                                                                  const int32_t* __restrict__ lengths, int nsamp, int T,
                                                                  FrontendTables tb, float* __restrict__ feat,
                                                                  int feat_rows) {
  extern __shared__ __align__(16) float smf[];
  Cx<float>* S = reinterpret_cast<Cx<float>*>(smf);            // 8 warps x 576
  Cx<float>* tw = S + 8 * kFftScratch;                         // 512
  Cx<float>* tw64 = tw + kNfft;                                // 72
  float* ybuf = reinterpret_cast<float*>(tw64 + kFftTw64);     // kChunk
  float* aux = ybuf + kChunk;                                  // mfcc: DCT rows (676) / stft: window (512)
  int* sbins = reinterpret_cast<int*>(aux + 704);              // 66
  float* sinvw = reinterpret_cast<float*>(sbins + 68);         // 65
  float* lm = sinvw + 68;                                      // 8 warps x 2 frames x 64 filters
  float* ostage = lm + 8 * 2 * kMaxFilt;                       // F x 17

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int tblk = blockIdx.x * kBlkFrames;
  const int F = tb.ncep;
  const int len = lengths ? max(0, min(lengths[b], nsamp)) : nsamp;
  int nfr = kStft ? 1 + len / kFrameStep
                  : (len <= kFrameLen ? 1 : 1 + (len - kFrameLen + kFrameStep - 1) / kFrameStep);
  nfr = min(nfr, T);

  for (int i = tid; i < kNfft; i += 256) tw[i] = g_fc.tw[i];
  if (tid < kFftTw64) tw64[tid] = g_fc.tw64[tid];
  if (kStft) {
    for (int i = tid; i < kNfft; i += 256) aux[i] = g_fc.hann[i];
  } else {
    if (tb.kind == 0)
      for (int i = tid; i < 26 * 26; i += 256) aux[i] = g_fc.dct[i];
    if (tid < tb.nfilt + 2) sbins[tid] = tb.bins[tid];
    if (tid < tb.nfilt + 1) sinvw[tid] = tb.inv_width[tid];
  }
  // ---- stage the block's samples.  Every thread issues ALL its global loads before it touches one of them (the loop
  // used to wait out one DRAM round trip per iteration: a third of the kernel's stall samples); the raw samples rest
  // in the FFT scratch, which is idle until the frames are built.
  const float* x = wav + (size_t)b * nsamp;
  float* rawb = reinterpret_cast<float*>(S);               // rawb[i] = x[c0 + i - 1]
  constexpr int kPer = (kChunk + 1 + 255) / 256;
  const int c0 = kStft ? tblk * kFrameStep - kNfft / 2 : tblk * kFrameStep;
  {
    float v[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = tid + 256 * j;
      int s = c0 + i - 1;
      if (kStft && tb.pad_mode == 0) {                     // librosa reflect padding (no edge repeat)
        if (s < 0) s = -s;
        if (s >= len) s = 2 * (len - 1) - s;
      }
      // int16 PCM: value / 32768 is exact in f32, so the two input formats give the same bits
      v[j] = (i <= kChunk && s >= 0 && s < len)
                 ? (tb.pcm16 ? (float)__ldg(reinterpret_cast<const int16_t*>(wav) + (size_t)b * nsamp + s) * (1.0f / 32768.0f)
                             : __ldg(x + s))
                 : 0.f;
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = tid + 256 * j;
      if (i <= kChunk) rawb[i] = v[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < kChunk; i += 256) {
    // stft (librosa.stft, center=True): frame t covers samples [160 t - 256, 160 t + 256) of the padded signal;
    // mfcc / fbank: pre-emphasis y[s] = x[s] - 0.97 x[s-1], y[0] = x[0] (rawb holds 0 for s - 1 < 0), zeros from `len` on
    ybuf[i] = kStft ? rawb[i + 1] : (c0 + i < len ? rawb[i + 1] - kPreemph * rawb[i] : 0.f);
  }
  __syncthreads();

  const int tl = 2 * warp;                 // this warp's frames: tblk + tl, tblk + tl + 1
  const bool live1 = tblk + tl < nfr, live2 = tblk + tl + 1 < nfr;
  if (!live1) {                            // padding frames of a ragged batch (or beyond T): zeros
    for (int f = lane; f < F; f += 32) {
      ostage[f * kOPitch + tl] = 0.f;
      ostage[f * kOPitch + tl + 1] = 0.f;
    }
  } else {
    // ---- two real frames ride one complex transform: z = x1 + i x2
    Cx<float> r[16];
    const float* y1 = ybuf + tl * kFrameStep;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      const int n = fft512_input_index(lane, s);
      float v1 = 0.f, v2 = 0.f;
      if (kStft) {
        const float w = aux[n];
        v1 = y1[n] * w;
        v2 = live2 ? y1[n + kFrameStep] * w : 0.f;
      } else if ((s & 7) < 6 || ((s & 7) == 6 && n < kFrameLen)) {      // zero pad 400 -> 512
        v1 = y1[n];
        v2 = live2 ? y1[n + kFrameStep] : 0.f;
      }
      r[s].x = v1;
      r[s].y = v2;
    }
    Cx<float>* Sw = S + warp * kFftScratch;
    fft512_pass1(lane, r, tw, Sw);
    __syncwarp();
    fft512_load2(lane, r, Sw);
    __syncwarp();
    fft512_pass2(lane, r, tw64, Sw);
    __syncwarp();
    fft512_load3(lane, r, Sw);
    __syncwarp();
    fft512_pass3(lane, r, Sw);
    __syncwarp();
    // ---- split the two spectra: X1[k] = (Z[k] + conj Z[N-k]) / 2, X2[k] = (Z[k] - conj Z[N-k]) / (2i); bins
    // k = lane + 32 i (i < 8) and k = 256 (lane 0)
    float p1[9], p2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const int k = lane + 32 * i;
      p1[i] = 0.f;
      p2[i] = 0.f;
      if (i < 8 || lane == 0) {
        const Cx<float> a = Sw[fft512_spec_index(k)], c = Sw[fft512_spec_index((kNfft - k) & (kNfft - 1))];
        const float x1r = 0.5f * (a.x + c.x), x1i = 0.5f * (a.y - c.y);
        const float x2r = 0.5f * (a.y + c.y), x2i = 0.5f * (c.x - a.x);
        p1[i] = x1r * x1r + x1i * x1i;
        p2[i] = x2r * x2r + x2i * x2i;
      }
    }
    if (kStft) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int k = lane + 32 * i;
        if ((i < 8 || lane == 0) && k < F) {
          ostage[k * kOPitch + tl] = log1pf(sqrtf(p1[i]));
          ostage[k * kOPitch + tl + 1] = live2 ? log1pf(sqrtf(p2[i])) : 0.f;
        }
      }
    } else {
      __syncwarp();                        // every lane has read the spectrum: reuse the scratch for the power bins
      float e1 = 0.f, e2 = 0.f;
      const float inv_n = 1.f / (float)kNfft;
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        if (i < 8 || lane == 0) {
          const float q1 = p1[i] * inv_n, q2 = p2[i] * inv_n;
          e1 += q1;
          e2 += q2;
          Sw[lane + 32 * i].x = q1;
          Sw[lane + 32 * i].y = q2;
        }
      }
      e1 = warp_sum(e1);
      e2 = warp_sum(e2);
      if (e1 == 0.f) e1 = 2.220446049250313e-16f;
      if (e2 == 0.f) e2 = 2.220446049250313e-16f;
      __syncwarp();
      // ---- mel filterbank: lane j <-> filter j (two rounds when nfilt > 32), both frames at once
      float* lm1 = lm + warp * (2 * kMaxFilt);
      float* lm2 = lm1 + kMaxFilt;
      for (int j = lane; j < tb.nfilt; j += 32) {
        const int b0 = sbins[j], b1 = sbins[j + 1], b2 = sbins[j + 2];
        const float up = sinvw[j], dn = sinvw[j + 1];
        float s1 = 0.f, s2 = 0.f;
        for (int i = b0; i < b1; ++i) {
          const float w = (float)(i - b0) * up;
          const Cx<float> pw = Sw[i];
          s1 = fmaf(pw.x, w, s1);
          s2 = fmaf(pw.y, w, s2);
        }
        for (int i = b1; i < b2; ++i) {
          const float w = (float)(b2 - i) * dn;
          const Cx<float> pw = Sw[i];
          s1 = fmaf(pw.x, w, s1);
          s2 = fmaf(pw.y, w, s2);
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
          const float* d = aux + n * 26;
#pragma unroll 2
          for (int j = 0; j < 26; ++j) {
            c1 = fmaf(d[j], lm1[j], c1);
            c2 = fmaf(d[j], lm2[j], c2);
          }
          if (n == 0) { c1 = logf(e1); c2 = logf(e2); }          // appendEnergy=True
          ostage[n * kOPitch + tl] = c1;
          ostage[n * kOPitch + tl + 1] = live2 ? c2 : 0.f;
        }
      } else {
        for (int f = lane; f < F; f += 32) {
          ostage[f * kOPitch + tl] = lm1[f];
          ostage[f * kOPitch + tl + 1] = live2 ? lm2[f] : 0.f;
        }
      }
    }
  }
  __syncthreads();
  float* out = feat + (size_t)b * feat_rows * T;          // feat_rows = F (1 + delta order): rows of one utterance
  for (int i = tid; i < F * kBlkFrames; i += 256) {
    const int f = i >> 4, tt = i & 15;
    if (tblk + tt < T) out[(size_t)f * T + tblk + tt] = ostage[f * kOPitch + tt];
  }
}

// grid (B, ceil(F/8)); block 256: one coefficient row per warp, held in shared memory (dynamic: 8 T floats) between the
// statistics and the two outputs -- f32 rows in place and 16-byte (8-channel) pieces of the channels-last bf16 copy the
// TDNN consumes.  Same summation order as frontend_cmvn_kernel (bitwise identical results).
__global__ void __launch_bounds__(256) frontend_cmvn2_kernel(float* __restrict__ feat, const int32_t* __restrict__ lengths,
                                                             int nsamp, int T, int F, int cmvn, int stft, int delta,
                                                             uint16_t* __restrict__ out_bf16, int ld) {
  extern __shared__ __align__(16) float rows[];      // [8][T]
  const int b = blockIdx.x, f0 = blockIdx.y * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lengths ? max(0, min(lengths[b], nsamp)) : nsamp;
  int nfr = stft ? 1 + len / kFrameStep : (len <= kFrameLen ? 1 : 1 + (len - kFrameLen + kFrameStep - 1) / kFrameStep);
  nfr = min(nfr, T);
  const int Fout = F * (1 + delta);                  // rows of one utterance: [feat | delta(feat, 1) | delta(feat, 2)]
  const int f = f0 + warp;
  float* srow = rows + warp * T;
  if (f < F) {
    float* row = feat + ((size_t)b * Fout + f) * T;
    float s = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float v = row[t];
      srow[t] = v;
      if (t < nfr) s += v;
    }
    float mean = 0.f, inv = 1.f;
    if (cmvn) {
      mean = warp_sum(s) / (float)nfr;
      float q = 0.f;
      for (int t = lane; t < nfr; t += 32) { const float d = srow[t] - mean; q = fmaf(d, d, q); }
      q = warp_sum(q) / (float)nfr;
      inv = 1.f / (sqrtf(q) + 2e-12f);
    }
    for (int t = lane; t < T; t += 32) {
      const float v = t < nfr ? (srow[t] - mean) * inv : 0.f;
      row[t] = v;
      srow[t] = v;
    }
  } else {
    for (int t = lane; t < T; t += 32) srow[t] = 0.f;
  }
  __syncthreads();
  uint16_t* ob = out_bf16 ? out_bf16 + (size_t)b * T * ld : nullptr;
  if (delta > 0 && f < F) {
    // python_speech_features.delta(feat, N) of the NORMALISED features, N = 1 and N = 2 (`_delta`, datasets.py:217-225):
    // sum_n n (f[t+n] - f[t-n]) / (2 sum_n n^2) with the edges repeated; the padding frames of a ragged batch stay zero
    for (int k = 1; k <= delta; ++k) {
      float* drow = feat + ((size_t)b * Fout + k * F + f) * T;
      for (int t = lane; t < T; t += 32) {
        float d = 0.f;
        if (t < nfr) {
          const float p1 = srow[min(t + 1, nfr - 1)], m1 = srow[max(t - 1, 0)];
          if (k == 1) d = (p1 - m1) * 0.5f;
          else d = ((p1 - m1) + 2.f * (srow[min(t + 2, nfr - 1)] - srow[max(t - 2, 0)])) * 0.1f;
        }
        drow[t] = d;
        if (ob) ob[(size_t)t * ld + k * F + f] = (uint16_t)(pack_bf16x2(d, 0.f) & 0xffffu);
      }
    }
  }
  if (ob) {
    if (delta == 0 || f0 + 8 <= F) {                 // 8 channels of one frame per 16-byte store
      for (int t = threadIdx.x; t < T; t += 256) {
        uint4 o;
        o.x = pack_bf16x2(rows[t], rows[T + t]);
        o.y = pack_bf16x2(rows[2 * T + t], rows[3 * T + t]);
        o.z = pack_bf16x2(rows[4 * T + t], rows[5 * T + t]);
        o.w = pack_bf16x2(rows[6 * T + t], rows[7 * T + t]);
        *reinterpret_cast<uint4*>(ob + (size_t)t * ld + f0) = o;
      }
    } else {                                         // partial last group next to the delta channels: element stores
      for (int i = threadIdx.x; i < T * (F - f0); i += 256) {
        const int t = i / (F - f0), r = i - t * (F - f0);
        ob[(size_t)t * ld + f0 + r] = (uint16_t)(pack_bf16x2(rows[r * T + t], 0.f) & 0xffffu);
      }
    }
    // zero the padded channels: whole groups of 8 from ceil8(Fout) on (with delta, also the tail of Fout's own group)
    const int g0 = (Fout + 7) >> 3, ng = (ld >> 3) - g0;
    for (int i = blockIdx.y * 256 + threadIdx.x; i < T * ng; i += 256 * gridDim.y) {
      const int t = i / ng, g = g0 + i - t * ng;
      *reinterpret_cast<uint4*>(ob + (size_t)t * ld + g * 8) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (delta > 0 && blockIdx.y == 0) {
      const int tail = min(ld, g0 * 8) - Fout;
      for (int i = threadIdx.x; i < T * tail; i += 256) {
        const int t = i / tail;
        ob[(size_t)t * ld + Fout + (i - t * tail)] = 0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host-side tables
inline double hz2mel(double hz) { return 2595.0 * log10(1.0 + hz / 700.0); }
inline double mel2hz(double mel) { return 700.0 * (pow(10.0, mel / 2595.0) - 1.0); }

// python_speech_features.get_filterbanks bin edges (nfft 512, 16 kHz, 0..8000 Hz) for kinds 0-2; nothing for stft.
inline void fill_frontend_tables(FrontendTables* tb, int kind, int F, int pad_mode) {
  const int nfilt = kind == 3 ? 0 : (kind == 0 ? 26 : F);
  *tb = FrontendTables{};
  tb->nfilt = nfilt; tb->ncep = F; tb->kind = kind; tb->pad_mode = pad_mode; tb->pcm16 = 0;
  if (nfilt == 0) return;
  const double lo = hz2mel(0.0), hi = hz2mel(8000.0);
  for (int i = 0; i < nfilt + 2; ++i) {
    const double mel = lo + (hi - lo) * (double)i / (double)(nfilt + 1);
    tb->bins[i] = (int)floor((kNfft + 1) * mel2hz(mel) / 16000.0);
  }
  for (int i = 0; i < nfilt + 1; ++i) {
    const int w = tb->bins[i + 1] - tb->bins[i];
    tb->inv_width[i] = w > 0 ? 1.0f / (float)w : 0.0f;
  }
}

inline void fill_frontend_const(FrontendConst* h) {
  const double pi = 3.14159265358979323846;
  for (int p = 0; p < kNfft; ++p) {
    const double a = -2.0 * pi * (double)p / (double)kNfft;
    h->tw[p].x = (float)cos(a);
    h->tw[p].y = (float)sin(a);
    // scipy.signal.get_window('hann', 400, fftbins=True), centred in the 512-point frame (librosa pad_center)
    const int w = p - (kNfft - kFrameLen) / 2;
    h->hann[p] = (w >= 0 && w < kFrameLen) ? (float)(0.5 - 0.5 * cos(2.0 * pi * (double)w / (double)kFrameLen)) : 0.f;
  }
  for (int i = 0; i < kFftTw64; ++i) h->tw64[i].x = h->tw64[i].y = 0.f;
  for (int m2 = 0; m2 < 8; ++m2)
    for (int j1 = 0; j1 < 8; ++j1) {
      const double a = -2.0 * pi * (double)(m2 * j1) / 64.0;
      h->tw64[fft512_tw64_index(m2, j1)].x = (float)cos(a);
      h->tw64[fft512_tw64_index(m2, j1)].y = (float)sin(a);
    }
  // scipy dct(type=2, norm='ortho') rows over the 26 log-mel energies, cepstral lifter (L = 22) folded in
  for (int n = 0; n < 26; ++n)
    for (int j = 0; j < 26; ++j) {
      const double ortho = n == 0 ? sqrt(1.0 / 26.0) : sqrt(2.0 / 26.0);
      const double lift = 1.0 + 11.0 * sin(pi * (double)n / 22.0);
      h->dct[n * 26 + j] = (float)(ortho * lift * cos(pi * (double)(n * (2 * j + 1)) / 52.0));
    }
}

}  // namespace dl
