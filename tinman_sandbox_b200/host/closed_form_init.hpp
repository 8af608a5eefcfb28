// closed_form_init.hpp — the reference's closed-form synthetic TestData with run-time nlev
// (compute_and_apply_rhs_test/cxx/pointers_only/data_structures.cpp:38-92,117-163), shared by the host-side drivers.
#ifndef CAAR_CLOSED_FORM_INIT_HPP
#define CAAR_CLOSED_FORM_INIT_HPP

#include <cmath>
#include <cstddef>
#include <vector>

#include "caar_b200.h"

namespace caar_host {

struct HostData {  // a TestData with run-time nlev
  caar_dims dims;
  std::vector<double> f[CAAR_NUM_FIELDS];
  caar_constants c;
  caar_control ctl;
  double dvv[16], ps0;
  std::vector<double> hyai;
  caar_arrays slice(int e0) const {  // pointers to element e0 of every array
    caar_arrays a;
    double** t = reinterpret_cast<double**>(&a);
    caar_dims one = dims;
    one.nelem = 1;
    for (int i = 0; i < CAAR_NUM_FIELDS; ++i)
      t[i] = const_cast<double*>(f[i].data()) + (size_t)e0 * caar_field_count(&one, i);
    return a;
  }
};

inline void init_data(HostData& d, int E, int L) {
  d.dims = {E, L, 4, 1, 3};
  for (int i = 0; i < CAAR_NUM_FIELDS; ++i) d.f[i].assign(caar_field_count(&d.dims, i), 0.0);
  enum { D, DINV, FCOR, MP, MET, RMET, DP, V, T, PHIS, QDP, ETA, OM, PHI, PEC, VN0 };
  for (int ie = 0; ie < E; ++ie)
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        const double e1 = ie + 1, i1 = i + 1, j1 = j + 1;
        const size_t q = (size_t)ie * 16 + i * 4 + j;
        d.f[FCOR][q] = std::sin(i1 + j1);
        d.f[MET][q] = i1 * j1;
        d.f[RMET][q] = 1. / d.f[MET][q];
        d.f[MP][q] = 2 * i1;
        d.f[PHIS][q] = i1 + j1;
        d.f[D][q * 4] = 1.0;
        d.f[D][q * 4 + 3] = 2.0;
        d.f[DINV][q * 4] = 1.0;
        d.f[DINV][q * 4 + 3] = 0.5;
        for (int k = 0; k < L; ++k) {
          const double k1 = k + 1;
          const size_t n = ((size_t)ie * L + k) * 16 + i * 4 + j;
          d.f[PHI][n] = std::cos(i1 + 3 * j1) + k1;
          d.f[VN0][2 * n] = d.f[VN0][2 * n + 1] = 1.0;
          d.f[PEC][n] = 1.0;
          d.f[OM][n] = j1 * j1;
          d.f[QDP][((size_t)ie * 2 * L + k) * 16 + i * 4 + j] = 1.0 + std::sin(i1 * j1 * k1);
          for (int t = 0; t < 3; ++t) {
            const double t1 = t + 1;
            const size_t m = (((size_t)ie * 3 + t) * L + k) * 16 + i * 4 + j;
            d.f[DP][m] = 10.0 * k1 + e1 + i1 + j1 + t1;
            d.f[V][2 * m] = 1.0 + 0.5 * k1 + i1 + j1 + 0.2 * e1 + 2.0 * t1;
            d.f[V][2 * m + 1] = 1.0 + 0.5 * k1 + i1 + j1 + 0.2 * e1 + 3.0 * t1;
            d.f[T][m] = 1000.0 - k1 - i1 - j1 + 0.1 * e1 + t1;
          }
        }
      }
  const double Rgas = 287.04, cp = 1005.0;
  d.c = {1.0 / 6.376e6, 1.0, cp, 461.5, Rgas, Rgas / cp};
  d.ctl = {0, E, 0, 1, 2, 0, 1.0};
  d.ps0 = 10.0;
  d.hyai.resize(L + 1);
  for (int i = 0; i <= L; ++i) d.hyai[i] = L + 1 - i;
  static const double lit[16] = {-3.0000000000000000, -0.80901699437494745, 0.30901699437494745,
                                 -0.50000000000000000, 4.0450849718747373,  0.00000000000000000,
                                 -1.11803398874989490, 1.54508497187473700, -1.5450849718747370,
                                 1.11803398874989490,  0.00000000000000000, -4.04508497187473730,
                                 0.5000000000000000,   -0.30901699437494745, 0.80901699437494745,
                                 3.000000000000000000};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) d.dvv[i * 4 + j] = lit[j * 4 + i];
}

}  // namespace caar_host
#endif
