// 512-point complex FFT for one warp, 16 points per lane in registers: three radix-8 passes (512 = 8 x 8 x 8) with two
// transposes through a 576-entry shared-memory scratch.  The phase functions are `__host__ __device__` and take the lane
// index explicitly, so tests/fft512_host_check.cu runs the very same index arithmetic on the CPU (32 lanes executed in
// turn between the points where the kernel has a __syncwarp) against a direct DFT.
//
// Decomposition (forward transform, W_N = exp(-2 pi i / N), input index n = 64 n1 + n2, n2 = 8 m1 + m2):
//   pass 1   A[k1][n2]     = W_512^(n2 k1) * sum_n1 z[64 n1 + n2] W_8^(n1 k1)          lane <-> n2 = lane, lane + 32
//   pass 2   B[k1][j1][m2] = W_64^(m2 j1)  * sum_m1 A[k1][8 m1 + m2] W_8^(m1 j1)       lane <-> (k1 = lane/8 (+4), m2 = lane%8)
//   pass 3   Z[k1 + 8 j1 + 64 j2] = sum_m2 B[k1][j1][m2] W_8^(m2 j2)                   lane <-> (k1 = lane/8 (+4), j1 = lane%8)
// Scratch layouts are padded so that every warp-wide access is (near) bank-conflict free for 8-byte elements:
//   after pass 1: S[72 k1 + n2]        after pass 2: S[72 k1 + 9 j1 + m2]        spectrum: S[k + k/8]
#pragma once
#if defined(__CUDACC__)
#define DL_HD __host__ __device__ __forceinline__
#else
#define DL_HD inline
#endif

namespace dl {

constexpr int kFftN = 512;
constexpr int kFftScratch = 576;     // complex elements of scratch per warp

template <typename T>
struct Cx {
  T x, y;
};
template <typename T> DL_HD Cx<T> cadd(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> DL_HD Cx<T> csub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> DL_HD Cx<T> cmul(Cx<T> a, Cx<T> b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T> DL_HD Cx<T> cmul_mi(Cx<T> a) { return {a.y, -a.x}; }     // a * (-i)

// In-place 8-point forward DFT of v[0..7] (natural order in, natural order out).
template <typename T>
DL_HD void dft8(Cx<T>* v) {
  const T h = (T)0.70710678118654752440;
  // 4-point DFTs of the even and of the odd inputs
  Cx<T> s0 = cadd(v[0], v[4]), s1 = csub(v[0], v[4]), s2 = cadd(v[2], v[6]), s3 = cmul_mi(csub(v[2], v[6]));
  const Cx<T> e0 = cadd(s0, s2), e2 = csub(s0, s2), e1 = cadd(s1, s3), e3 = csub(s1, s3);
  s0 = cadd(v[1], v[5]); s1 = csub(v[1], v[5]); s2 = cadd(v[3], v[7]); s3 = cmul_mi(csub(v[3], v[7]));
  const Cx<T> o0 = cadd(s0, s2), o2 = cmul_mi(csub(s0, s2));
  Cx<T> o1 = cadd(s1, s3), o3 = csub(s1, s3);
  o1 = {h * (o1.x + o1.y), h * (o1.y - o1.x)};        // * W_8^1 = (1 - i) / sqrt 2
  o3 = {h * (o3.y - o3.x), -h * (o3.x + o3.y)};       // * W_8^3 = (-1 - i) / sqrt 2
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// Input element held in r[8 b + n1] of `lane` before pass 1:  n = 64 n1 + lane + 32 b.
DL_HD int fft512_input_index(int lane, int slot) { return 64 * (slot & 7) + lane + 32 * (slot >> 3); }

// tw[p] = exp(-2 pi i p / 512), p in [0, 512)
template <typename T>
DL_HD void fft512_pass1(int lane, Cx<T>* r, const Cx<T>* tw, Cx<T>* S) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < 2; ++b) {
    const int n2 = lane + 32 * b;
    dft8(r + 8 * b);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k1 = 0; k1 < 8; ++k1) {
      const Cx<T> a = k1 == 0 ? r[8 * b] : cmul(r[8 * b + k1], tw[n2 * k1]);
      S[72 * k1 + n2] = a;
    }
  }
}

template <typename T>
DL_HD void fft512_load2(int lane, Cx<T>* r, const Cx<T>* S) {
  const int m2 = lane & 7;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < 2; ++b) {
    const int k1 = (lane >> 3) + 4 * b;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m1 = 0; m1 < 8; ++m1) r[8 * b + m1] = S[72 * k1 + 8 * m1 + m2];
  }
}

// tw64[9 m2 + j1] = exp(-2 pi i m2 j1 / 64): compact copy of the pass-2 twiddles, row pitch 9 so that the eight
// distinct addresses of a warp-wide read fall into different banks.
DL_HD int fft512_tw64_index(int m2, int j1) { return 9 * m2 + j1; }
constexpr int kFftTw64 = 72;

template <typename T>
DL_HD void fft512_pass2(int lane, Cx<T>* r, const Cx<T>* tw64, Cx<T>* S) {
  const int m2 = lane & 7;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < 2; ++b) {
    const int k1 = (lane >> 3) + 4 * b;
    dft8(r + 8 * b);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j1 = 0; j1 < 8; ++j1) {
      const Cx<T> a = j1 == 0 ? r[8 * b] : cmul(r[8 * b + j1], tw64[fft512_tw64_index(m2, j1)]);
      S[72 * k1 + 9 * j1 + m2] = a;
    }
  }
}

template <typename T>
DL_HD void fft512_load3(int lane, Cx<T>* r, const Cx<T>* S) {
  const int j1 = lane & 7;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < 2; ++b) {
    const int k1 = (lane >> 3) + 4 * b;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m2 = 0; m2 < 8; ++m2) r[8 * b + m2] = S[72 * k1 + 9 * j1 + m2];
  }
}

DL_HD int fft512_spec_index(int k) { return k + (k >> 3); }

template <typename T>
DL_HD void fft512_pass3(int lane, Cx<T>* r, Cx<T>* S) {
  const int j1 = lane & 7;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < 2; ++b) {
    const int k1 = (lane >> 3) + 4 * b;
    dft8(r + 8 * b);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j2 = 0; j2 < 8; ++j2) S[fft512_spec_index(k1 + 8 * j1 + 64 * j2)] = r[8 * b + j2];
  }
}

}  // namespace dl
