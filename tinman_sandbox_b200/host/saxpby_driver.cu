// saxpby_driver.cu — the reference's bandwidth calibrator (saxpby_test/cxx/main.cpp:12-55) on a B200: same command
// line (optional I1), same sizes (I1 x 128 x 256 doubles per array, common.hpp:9-11), same initial values, the same
// 100 sweeps of x = 3x + 5y (main.cpp:39-41, common.cpp:3-15), the same two "name: seconds s" timer lines — with the
// two arrays resident in HBM and the sweep done by the library's kernel (caar_saxpby_device). Adds the achieved GB/s
// (24 bytes per element per sweep) and, with --check, compares x against the same recurrence on the host.
//   saxpby_driver [I1=1000] [--check]
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "caar_b200.h"

namespace {
constexpr int I2 = 128, I3 = 256;
double now() { return std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count(); }
}  // namespace

int main(int argc, char* argv[]) {
  int I1 = 1000;
  bool check = false;
  for (int i = 1; i < argc; ++i) {
    if (std::strcmp(argv[i], "--check") == 0) check = true;
    else I1 = std::atoi(argv[i]);
  }
  if (I1 < 1 || caar_device_count() < 1) {
    std::fprintf(stderr, "saxpby_driver: need I1 >= 1 and a CUDA device (%s); there is no CPU path\n", caar_last_error());
    return 2;
  }
  const size_t n = (size_t)I1 * I2 * I3;
  std::vector<double> x(n), y(n);
  std::cout << I1 << "    " << I2 << "    " << I3 << std::endl;
  double *dx = nullptr, *dy = nullptr;
  {
    const double t0 = now();
    for (int i = 0; i < I1; ++i)
      for (int j = 0; j < I2; ++j)
        for (int k = 0; k < I3; ++k) {
          const size_t q = (size_t)k + (size_t)I3 * j + (size_t)I2 * I3 * i;
          x[q] = (double)i * j * k;  // the reference multiplies ints (main.cpp:30-31) and overflows for large I1;
          y[q] = (double)i * i * j * j * k * k;  // doubles here
        }
    if (cudaMalloc(&dx, n * sizeof(double)) != cudaSuccess || cudaMalloc(&dy, n * sizeof(double)) != cudaSuccess) {
      std::fprintf(stderr, "saxpby_driver: cudaMalloc of 2 x %zu bytes failed\n", n * sizeof(double));
      return 2;
    }
    cudaMemcpy(dx, x.data(), n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(dy, y.data(), n * sizeof(double), cudaMemcpyHostToDevice);
    std::cout << "Init: " << now() - t0 << " s" << std::endl;
  }
  const int sweeps = 100;
  // a = 3, b = 5 overflow to inf after ~100 sweeps of growth 3^100 only for huge values; fine for a bandwidth test,
  // exactly like the reference
  cudaDeviceSynchronize();
  const double t0 = now();
  for (int it = 0; it < sweeps; ++it)
    if (int rc = caar_saxpby_device(3.0, 5.0, dx, dy, n, nullptr)) {
      std::fprintf(stderr, "caar_saxpby_device failed (%d): %s\n", rc, caar_last_error());
      return 2;
    }
  cudaDeviceSynchronize();
  const double sec = now() - t0;
  std::cout << "saxpby: " << sec << " s" << std::endl;
  std::printf("   ---> %.1f GB/s (24 bytes per element per sweep, %d sweeps, %zu elements)\n",
              24.0 * n * sweeps / sec / 1e9, sweeps, n);
  int bad = 0;
  if (check) {
    std::vector<double> got(n);
    cudaMemcpy(got.data(), dx, n * sizeof(double), cudaMemcpyDeviceToHost);
    for (size_t q = 0; q < n; q += 977) {
      double w = x[q];
      for (int it = 0; it < sweeps; ++it) w = 3.0 * w + 5.0 * y[q];
      const double err = std::fabs(got[q] - w) / (std::fabs(w) > 0 ? std::fabs(w) : 1.0);
      if (!(err <= 1e-13) && !(std::isinf(w) && std::isinf(got[q]))) ++bad;
    }
    std::printf("   ---> check: %s\n", bad ? "MISMATCH" : "ok");
  }
  cudaFree(dx);
  cudaFree(dy);
  return bad ? 1 : 0;
}
