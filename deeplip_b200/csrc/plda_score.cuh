// S5: PLDA trial scoring, the GPU form of `eer_plda_grid` / `eer_plda_lomgrid`
// (models/audio_models/utils.py:285-329): per trial the reference transforms the two embeddings D -> U_model with the
// fitted `plda` model and takes `calc_same_diff_log_likelihood_ratio`.  Here every utterance is transformed ONCE
// (an affine map R x D, R <= 32 relevant dimensions) and a trial is 2 R fused multiply-adds on the two R-vectors:
//   llr = c0 + sum_r k1[r] (a_r + b_r)^2 - k2[r] (a_r^2 + b_r^2),
//   k1 = psi / (2 (2 psi + 1)),  k2 = psi / (2 (psi + 1)),  c0 = sum_r log(psi + 1) - log(2 psi + 1) / 2
// which is logp({a,b}) - logp({a}) - logp({b}) of the package's marginal likelihood written out (deeplip_b200/plda.py).
// Free of CUDA-runtime dependencies (tests/frontend_cpu_emul.cpp runs this source on CPU threads).
// The includer provides: warp_sum(float), __ldg.
#pragma once
#include <stdint.h>

namespace dl {

constexpr int kPldaMaxR = 32;

// grid (ceil(n_utt / 8)); block 256: one warp per utterance.  u[row, r] = bias[r] + sum_d emb[row, d] M[r, d]
__global__ void __launch_bounds__(256) plda_transform_kernel(const float* __restrict__ emb, int n_utt, int D,
                                                             const float* __restrict__ M, const float* __restrict__ bias,
                                                             int R, float* __restrict__ u) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_utt) return;
  float acc[kPldaMaxR];
#pragma unroll
  for (int r = 0; r < kPldaMaxR; ++r) acc[r] = 0.f;
  const float* x = emb + (size_t)row * D;
  for (int d = lane; d < D; d += 32) {
    const float xv = __ldg(x + d);
#pragma unroll
    for (int r = 0; r < kPldaMaxR; ++r)
      if (r < R) acc[r] = fmaf(xv, __ldg(M + (size_t)r * D + d), acc[r]);
  }
  float mine = 0.f;
#pragma unroll
  for (int r = 0; r < kPldaMaxR; ++r) {
    if (r < R) {                       // warp-uniform
      const float s = warp_sum(acc[r]);
      if (lane == r) mine = s;
    }
  }
  if (lane < R) u[(size_t)row * R + lane] = mine + __ldg(bias + lane);
}

// one thread per trial
__global__ void __launch_bounds__(256) plda_llr_trials_kernel(const float* __restrict__ u, int n_utt, int R,
                                                              const float* __restrict__ k1, const float* __restrict__ k2,
                                                              float c0, const int32_t* __restrict__ enrol,
                                                              const int32_t* __restrict__ test, int n_trials,
                                                              float* __restrict__ scores) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= n_trials) return;
  const int i = enrol[t], j = test[t];
  if (i < 0 || i >= n_utt || j < 0 || j >= n_utt) {
    uint32_t nan_bits = 0x7fc00000u;
    scores[t] = *reinterpret_cast<float*>(&nan_bits);
    return;
  }
  const float* a = u + (size_t)i * R;
  const float* b = u + (size_t)j * R;
  float s = c0;
  for (int r = 0; r < R; ++r) {
    const float av = a[r], bv = b[r], sum = av + bv;
    s = fmaf(__ldg(k1 + r), sum * sum, s);
    s = fmaf(-__ldg(k2 + r), fmaf(av, av, bv * bv), s);
  }
  scores[t] = s;
}

}  // namespace dl
