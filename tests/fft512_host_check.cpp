// CPU execution of the warp FFT's phase functions (deeplip_b200/csrc/fft512.cuh): the 32 lanes run in turn between the
// points where the kernel has a __syncwarp; the result is compared with a direct O(N^2) DFT in double precision.
// Built and run by tests/test_host_logic.py::test_fft512_phase_functions_on_cpu.  Prints "max_err <float> <double>".
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../deeplip_b200/csrc/fft512.cuh"

template <typename T>
double run(unsigned seed) {
  using namespace dl;
  std::vector<Cx<double>> z(kFftN), ref(kFftN);
  srand(seed);
  for (auto& v : z) v = {rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
  for (int k = 0; k < kFftN; ++k) {
    Cx<double> acc = {0, 0};
    for (int n = 0; n < kFftN; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)k * n % kFftN) / kFftN;
      acc = cadd(acc, cmul(z[n], Cx<double>{cos(a), sin(a)}));
    }
    ref[k] = acc;
  }
  std::vector<Cx<T>> tw(kFftN), S(kFftScratch, Cx<T>{(T)1e30, (T)1e30});
  for (int p = 0; p < kFftN; ++p) tw[p] = {(T)cos(-2.0 * M_PI * p / kFftN), (T)sin(-2.0 * M_PI * p / kFftN)};
  std::vector<Cx<T>> tw64(kFftTw64, Cx<T>{(T)1e30, (T)1e30});
  for (int m2 = 0; m2 < 8; ++m2)
    for (int j1 = 0; j1 < 8; ++j1)
      tw64[fft512_tw64_index(m2, j1)] = {(T)cos(-2.0 * M_PI * m2 * j1 / 64), (T)sin(-2.0 * M_PI * m2 * j1 / 64)};
  Cx<T> r[32][16];
  for (int lane = 0; lane < 32; ++lane)
    for (int s = 0; s < 16; ++s) {
      const Cx<double> v = z[fft512_input_index(lane, s)];
      r[lane][s] = {(T)v.x, (T)v.y};
    }
  for (int lane = 0; lane < 32; ++lane) fft512_pass1(lane, r[lane], tw.data(), S.data());
  for (int lane = 0; lane < 32; ++lane) fft512_load2(lane, r[lane], S.data());
  for (int lane = 0; lane < 32; ++lane) fft512_pass2(lane, r[lane], tw64.data(), S.data());
  for (int lane = 0; lane < 32; ++lane) fft512_load3(lane, r[lane], S.data());
  for (int lane = 0; lane < 32; ++lane) fft512_pass3(lane, r[lane], S.data());
  double err = 0;
  for (int k = 0; k < kFftN; ++k) {
    const Cx<T> g = S[fft512_spec_index(k)];
    err = fmax(err, fmax(fabs((double)g.x - ref[k].x), fabs((double)g.y - ref[k].y)));
  }
  return err;
}

int main() {
  double ef = 0, ed = 0;
  for (unsigned s = 1; s <= 3; ++s) { ef = fmax(ef, run<float>(s)); ed = fmax(ed, run<double>(s)); }
  printf("max_err %.3e %.3e\n", ef, ed);
  return (ef < 2e-4 && ed < 1e-11) ? 0 : 1;
}
