// oracle/ref_saxpby_capi.cpp — TEST INFRASTRUCTURE ONLY.
// extern "C" entry to the reference's own saxpby (saxpby_test/cxx/common.cpp:3-15), which is compiled
// from /root/reference together with this file into oracle/_ref/libsaxpby_ref.so (see oracle/Makefile).
// The reference sizes its loop with the global `I1` (saxpby_test/cxx/main.cpp:12) and the constexpr
// I2=128, I3=256 (common.hpp:9-11): n = I1*128*256 doubles.
#include "common.hpp"  // reference header (via -I)

#include <chrono>

int I1 = 1;

extern "C" double saxpby_ref_run(double a, double b, double* x, const double* y, int i1, int sweeps) {
  I1 = i1;
  auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < sweeps; ++s) saxpby(a, b, x, y);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}
