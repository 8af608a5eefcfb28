// oracle/ref_capi.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A thin extern "C" wrapper that is compiled TOGETHER WITH the reference's own, unmodified
// sources where they lie under /root/reference (compute_and_apply_rhs_test/cxx/pointers_only/
// {compute_and_apply_rhs,sphere_operators,data_structures}.cpp) into oracle/_ref/libcaar_ref_L<PLEV>.so
// by oracle/Makefile. It lets the tests and bench.py's cpu_baseline / --impl reference legs call the real
// Homme::compute_and_apply_rhs(TestData&) on caller-owned arrays.
//
// Nothing here restates reference arithmetic: it only builds a Homme::TestData that points at the
// caller's arrays and calls the reference entry points
//   Homme::compute_and_apply_rhs      (PO/compute_and_apply_rhs.hpp:9)
//   Homme::TestData::init_data        (PO/data_structures.cpp:165-172)
//   Homme::compute_norm               (PO/compute_and_apply_rhs.cpp:354-370)
// The reference headers are found through -I at build time; no reference source is copied here.

#include "data_structures.hpp"          // reference header (via -I)
#include "compute_and_apply_rhs.hpp"    // reference header (via -I)
#include "sphere_operators.hpp"         // reference header (via -I)

#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace Homme {
int num_elems = 10;  // the reference defines this global in its main.cpp (PO/main.cpp:9-12)
}

namespace {

struct RefArrays {  // same order as Homme::Arrays / caar_arrays
  double* p[16];
};

size_t field_count(int f, int E) {
  using namespace Homme;
  const size_t e = (size_t)E, pts = np * np;
  switch (f) {
    case 0: case 1: return e * pts * 4;
    case 2: case 3: case 4: case 5: case 9: return e * pts;
    case 6: case 8: return e * timelevels * nlev * pts;
    case 7: return e * timelevels * nlev * pts * 2;
    case 10: return e * qsize_d * 2 * nlev * pts;
    case 11: return e * nlevp * pts;
    case 12: case 13: case 14: return e * nlev * pts;
    case 15: return e * nlev * pts * 2;
  }
  return 0;
}

double** member(Homme::Arrays& a, int f) {
  switch (f) {
    case 0: return &a.elem_D;
    case 1: return &a.elem_Dinv;
    case 2: return &a.elem_fcor;
    case 3: return &a.elem_spheremp;
    case 4: return &a.elem_metdet;
    case 5: return &a.elem_rmetdet;
    case 6: return &a.elem_state_dp3d;
    case 7: return &a.elem_state_v;
    case 8: return &a.elem_state_T;
    case 9: return &a.elem_state_phis;
    case 10: return &a.elem_state_Qdp;
    case 11: return &a.elem_derived_eta_dot_dpdn;
    case 12: return &a.elem_derived_omega_p;
    case 13: return &a.elem_derived_phi;
    case 14: return &a.elem_derived_pecnd;
    case 15: return &a.elem_derived_vn0;
  }
  return nullptr;
}

void fill(Homme::TestData& d, double* const* arrays, const int* ctl7i, double dt2, const double* consts6,
          const double* dvv16, double ps0, const double* hyai) {
  for (int f = 0; f < 16; ++f) *member(d.arrays, f) = arrays[f];
  d.control.nets = ctl7i[0];
  d.control.nete = ctl7i[1];
  d.control.n0 = ctl7i[2];
  d.control.np1 = ctl7i[3];
  d.control.nm1 = ctl7i[4];
  d.control.qn0 = ctl7i[5];
  d.control.dt2 = dt2;
  d.constants.rrearth = consts6[0];
  d.constants.eta_ave_w = consts6[1];
  d.constants.cp = consts6[2];
  d.constants.Rwater_vapor = consts6[3];
  d.constants.Rgas = consts6[4];
  d.constants.kappa = consts6[5];
  for (int i = 0; i < Homme::np; ++i)
    for (int j = 0; j < Homme::np; ++j) d.deriv.Dvv[i][j] = dvv16[i * Homme::np + j];
  d.hvcoord.ps0 = ps0;
  for (int i = 0; i < Homme::nlevp; ++i) d.hvcoord.hyai[i] = hyai[i];
}

}  // namespace

extern "C" {

int caar_ref_nlev(void) { return Homme::nlev; }
int caar_ref_np(void) { return Homme::np; }
int caar_ref_qsize_d(void) { return Homme::qsize_d; }
int caar_ref_ntl(void) { return Homme::timelevels; }

size_t caar_ref_field_count(int field, int nelem) { return field_count(field, nelem); }

// Run the reference's own TestData::init_data() for `nelem` elements and copy every array, the
// constants (6), Dvv (16, row-major [i][j]), ps0 and hyai (nlev+1) out to the caller.
void caar_ref_init(int nelem, double* const* arrays, int* ctl6i, double* dt2, double* consts6, double* dvv16,
                   double* ps0, double* hyai) {
  Homme::num_elems = nelem;
  Homme::TestData d;
  d.init_data();
  for (int f = 0; f < 16; ++f)
    std::memcpy(arrays[f], *member(d.arrays, f), field_count(f, nelem) * sizeof(double));
  ctl6i[0] = d.control.nets;
  ctl6i[1] = d.control.nete;
  ctl6i[2] = d.control.n0;
  ctl6i[3] = d.control.np1;
  ctl6i[4] = d.control.nm1;
  ctl6i[5] = d.control.qn0;
  *dt2 = d.control.dt2;
  consts6[0] = d.constants.rrearth;
  consts6[1] = d.constants.eta_ave_w;
  consts6[2] = d.constants.cp;
  consts6[3] = d.constants.Rwater_vapor;
  consts6[4] = d.constants.Rgas;
  consts6[5] = d.constants.kappa;
  for (int i = 0; i < Homme::np; ++i)
    for (int j = 0; j < Homme::np; ++j) dvv16[i * Homme::np + j] = d.deriv.Dvv[i][j];
  *ps0 = d.hvcoord.ps0;
  for (int i = 0; i < Homme::nlevp; ++i) hyai[i] = d.hvcoord.hyai[i];
  d.cleanup_data();
}

// ncalls calls of the reference routine on caller-owned arrays; elements [nets,nete) are split into
// `nthreads` contiguous ranges, one std::thread each running the UNMODIFIED routine on its own
// TestData view (the reference's own nets/nete partition hook, PO/compute_and_apply_rhs.cpp:65-74).
// Returns wall seconds (steady_clock) spent in the calls.
double caar_ref_run(double* const* arrays, const int* ctl6i, double dt2, const double* consts6,
                    const double* dvv16, double ps0, const double* hyai, int ncalls, int nthreads) {
  const int nets = ctl6i[0], nete = ctl6i[1];
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nete - nets) nthreads = nete - nets > 0 ? nete - nets : 1;
  std::vector<Homme::TestData> views(nthreads);
  for (int t = 0; t < nthreads; ++t) {
    fill(views[t], arrays, ctl6i, dt2, consts6, dvv16, ps0, hyai);
    const long long n = nete - nets;
    views[t].control.nets = nets + (int)(n * t / nthreads);
    views[t].control.nete = nets + (int)(n * (t + 1) / nthreads);
  }
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1) {
    for (int c = 0; c < ncalls; ++c) Homme::compute_and_apply_rhs(views[0]);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&views, t, ncalls]() {
        for (int c = 0; c < ncalls; ++c) Homme::compute_and_apply_rhs(views[t]);
      });
    for (auto& th : pool) th.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// The three numbers print_results_2norm prints (PO/compute_and_apply_rhs.cpp:372-399), computed with the
// reference's own compute_norm (Kahan) per element, squared and summed, then sqrt.
void caar_ref_norms(double* const* arrays, int nets, int nete, int tl, double out3[3]) {
  using namespace Homme;
  double vn = 0, tn = 0, dn = 0;
  const size_t lev_pts = (size_t)nlev * np * np;
  for (int ie = nets; ie < nete; ++ie) {
    const double* v = arrays[7] + ((size_t)ie * timelevels + tl) * lev_pts * 2;
    const double* T = arrays[8] + ((size_t)ie * timelevels + tl) * lev_pts;
    const double* dp = arrays[6] + ((size_t)ie * timelevels + tl) * lev_pts;
    vn += std::pow(compute_norm(v, nlev * np * np * 2), 2);
    tn += std::pow(compute_norm(T, nlev * np * np), 2);
    dn += std::pow(compute_norm(dp, nlev * np * np), 2);
  }
  out3[0] = std::sqrt(vn);
  out3[1] = std::sqrt(tn);
  out3[2] = std::sqrt(dn);
}

// The reference's own divergence_sphere (PO/sphere_operators.cpp:50-89) on one 4x4 level `v` [4][4][2] of
// element `ie` of caller-owned geometry arrays; used to pin the operator the tracer-step oracle is built on.
void caar_ref_divergence_sphere(const double* v, double* const* arrays, int ie, const double* dvv16, double rrearth,
                                double* div) {
  Homme::TestData d;
  for (int f = 0; f < 16; ++f) *member(d.arrays, f) = arrays[f];
  d.constants.rrearth = rrearth;
  for (int i = 0; i < Homme::np; ++i)
    for (int j = 0; j < Homme::np; ++j) d.deriv.Dvv[i][j] = dvv16[i * Homme::np + j];
  Homme::divergence_sphere(v, d, ie, div);
}

}  // extern "C"
