// homme_shim.cpp — the drop-in: Homme::compute_and_apply_rhs(TestData&) on a B200.
//
// This translation unit REPLACES the reference's compute_and_apply_rhs.cpp and sphere_operators.cpp in the
// cxx test driver (compute_and_apply_rhs_test/cxx/pointers_only/): it is compiled against the reference's
// OWN headers (data_structures.hpp, compute_and_apply_rhs.hpp — found through -I, nothing is copied) and
// linked with the reference's unmodified main.cpp, data_structures.cpp and timer.cpp plus libcaar_b200.so.
// It defines every symbol main.cpp uses from the replaced file:
//
//   Homme::compute_and_apply_rhs(TestData&)   PO/compute_and_apply_rhs.hpp:9    -> C-ABI, CUDA kernels
//   Homme::compute_norm                       PO/compute_and_apply_rhs.cpp:354  (host, compensated sum)
//   Homme::print_results_2norm                PO/compute_and_apply_rhs.cpp:372  (same text format)
//   Homme::dump_results_to_file               PO/compute_and_apply_rhs.cpp:401  (same four files)
//
// Semantics: the host arrays are the truth, exactly as in the reference. Each call streams the slices the
// routine reads to the GPU, runs one RHS evaluation and streams back the slices it writes (caar_run_host:
// copy-in | kernel | copy-out pipelined over element chunks), so the driver's later reads (norms, dumps) see
// the results. A handle is created on first use and kept for the process
// (device mirrors are reused between calls); the caller's arrays are page-locked in place on first use
// (CAAR_PIN=0 disables) so the copies run at PCIe speed. Errors abort with a message, like the reference's
// own failure paths (std::abort, PO/compute_and_apply_rhs.cpp:414-445).
//
// Environment: CAAR_MODE=fast|strict (default fast), CAAR_DEVICE=<ordinal> (default 0), CAAR_PIN=0|1.
#include "compute_and_apply_rhs.hpp"  // reference header (via -I)
#include "data_structures.hpp"        // reference header (via -I)
#include "dimensions.hpp"             // reference header (via -I)

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>

#include "caar_b200.h"

namespace Homme {

extern int num_elems;  // defined by the driver (PO/main.cpp:9-12), sized the arrays (PO/data_structures.cpp:10)

namespace {

// One per process, allocated on first use and deliberately NEVER destroyed: no CUDA call may run during static
// destruction (the runtime may already be gone), so the handle is left to process exit. The page-locking of the
// caller's arrays, on the other hand, must end BEFORE the caller frees them (the reference driver delete[]s its arrays
// in cleanup_data, PO/main.cpp:140): print_results_2norm — which the driver calls right after its timed loop,
// PO/main.cpp:131 — releases it; a later compute_and_apply_rhs call simply pins again.
struct Session {
  caar_handle h = nullptr;
  int nelem = 0;
  int mode = CAAR_MODE_FAST;
  const double* pinned_key = nullptr;
  caar_arrays pinned = {};
  bool pin = true;
  long calls = 0;

  void unpin() {
    double* const* tab = reinterpret_cast<double* const*>(&pinned);
    for (int f = 0; f < CAAR_NUM_FIELDS; ++f)
      if (tab[f]) caar_host_unregister(tab[f]);
    pinned_key = nullptr;
    pinned = caar_arrays{};
  }
};

[[noreturn]] void die(const char* what, int rc) {
  std::fprintf(stderr, "caar_b200 shim: %s failed (code %d): %s\n", what, rc, caar_last_error());
  std::abort();
}

caar_arrays view(const Arrays& a) {
  caar_arrays v;
  v.elem_D = a.elem_D;
  v.elem_Dinv = a.elem_Dinv;
  v.elem_fcor = a.elem_fcor;
  v.elem_spheremp = a.elem_spheremp;
  v.elem_metdet = a.elem_metdet;
  v.elem_rmetdet = a.elem_rmetdet;
  v.elem_state_dp3d = a.elem_state_dp3d;
  v.elem_state_v = a.elem_state_v;
  v.elem_state_T = a.elem_state_T;
  v.elem_state_phis = a.elem_state_phis;
  v.elem_state_Qdp = a.elem_state_Qdp;
  v.elem_derived_eta_dot_dpdn = a.elem_derived_eta_dot_dpdn;
  v.elem_derived_omega_p = a.elem_derived_omega_p;
  v.elem_derived_phi = a.elem_derived_phi;
  v.elem_derived_pecnd = a.elem_derived_pecnd;
  v.elem_derived_vn0 = a.elem_derived_vn0;
  return v;
}

Session* g_session = nullptr;

Session& session(const TestData& data) {
  if (!g_session) g_session = new Session();  // leaked on purpose, see above
  Session& s = *g_session;
  if (s.h && s.nelem != num_elems) {
    if (s.pinned_key) s.unpin();
    caar_destroy(s.h);
    s.h = nullptr;
  }
  if (!s.h) {
    const char* m = std::getenv("CAAR_MODE");
    s.mode = (m && std::strcmp(m, "strict") == 0) ? CAAR_MODE_STRICT : CAAR_MODE_FAST;
    const char* d = std::getenv("CAAR_DEVICE");
    const char* p = std::getenv("CAAR_PIN");
    s.pin = !(p && std::strcmp(p, "0") == 0);
    caar_dims dims = {num_elems, nlev, np, qsize_d, timelevels};
    if (int rc = caar_create(&s.h, &dims, d ? std::atoi(d) : 0)) die("caar_create", rc);
    s.nelem = num_elems;
  }
  if (s.pin && s.pinned_key != data.arrays.elem_state_v) {
    if (s.pinned_key) s.unpin();
    caar_dims dims = {num_elems, nlev, np, qsize_d, timelevels};
    const caar_arrays v = view(data.arrays);
    double* const* tab = reinterpret_cast<double* const*>(&v);
    double** keep = reinterpret_cast<double**>(&s.pinned);
    for (int f = 0; f < CAAR_NUM_FIELDS; ++f)
      if (caar_host_register(tab[f], caar_field_count(&dims, f) * sizeof(double)) == CAAR_OK) keep[f] = tab[f];
    s.pinned_key = data.arrays.elem_state_v;
  }
  return s;
}

}  // namespace

void compute_and_apply_rhs(TestData& data) {
  Session& s = session(data);
  caar_constants c = {data.constants.rrearth, data.constants.eta_ave_w, data.constants.cp,
                      data.constants.Rwater_vapor, data.constants.Rgas, data.constants.kappa};
  if (int rc = caar_set_params(s.h, &c, &data.deriv.Dvv[0][0], data.hvcoord.ps0, data.hvcoord.hyai))
    die("caar_set_params", rc);
  caar_control ctl = {data.control.nets, data.control.nete, data.control.n0, data.control.np1,
                      data.control.nm1, data.control.qn0, data.control.dt2};
  const caar_arrays host = view(data.arrays);
  if (int rc = caar_run_host(s.h, &host, &ctl, s.mode, 0)) die("caar_run_host", rc);
  ++s.calls;
}

// sqrt of a compensated (Kahan) sum of squares — what the driver's norm check is built on
real compute_norm(const real* const field, int length) {
  real sum = 0, lost = 0;
  for (int n = 0; n < length; ++n) {
    const real term = field[n] * field[n] - lost;
    const real next = sum + term;
    lost = (next - sum) - term;
    sum = next;
  }
  return std::sqrt(sum);
}

void print_results_2norm(const TestData& data) {
  // the timed loop is over (PO/main.cpp:113-131): give the caller's arrays back unpinned before it can free them
  if (g_session && g_session->calls > 0 && g_session->pinned_key) g_session->unpin();
  const int tl = data.control.np1;
  const std::size_t lev_pts = static_cast<std::size_t>(nlev) * np * np;
  real acc[3] = {0, 0, 0};
  for (int ie = data.control.nets; ie < data.control.nete; ++ie) {
    const std::size_t slab = static_cast<std::size_t>(ie) * timelevels + tl;
    acc[0] += std::pow(compute_norm(data.arrays.elem_state_v + slab * lev_pts * 2, nlev * np * np * 2), 2);
    acc[1] += std::pow(compute_norm(data.arrays.elem_state_T + slab * lev_pts, nlev * np * np), 2);
    acc[2] += std::pow(compute_norm(data.arrays.elem_state_dp3d + slab * lev_pts, nlev * np * np), 2);
  }
  static const char* const label[3] = {"||v||_2  = ", "||T||_2  = ", "||dp||_2 = "};
  std::cout << "   ---> Norms:\n";
  for (int q = 0; q < 3; ++q)
    std::cout << "          " << label[q] << std::setprecision(17) << std::sqrt(acc[q]) << "\n";
}

void dump_results_to_file(const TestData& data) {
  static const char* const names[4] = {"elem_state_vx.txt", "elem_state_vy.txt", "elem_state_t.txt",
                                       "elem_state_dp3d.txt"};
  std::ofstream out[4];
  for (int f = 0; f < 4; ++f) {
    out[f].open(names[f]);
    if (!out[f].is_open()) {
      std::cout << "Error! Cannot open '" << names[f] << "'.\n";
      std::abort();
    }
    out[f].precision(6);
  }
  const int tl = data.control.np1;
  const std::size_t lev_pts = static_cast<std::size_t>(nlev) * np * np;
  for (int ie = data.control.nets; ie < data.control.nete; ++ie) {
    const std::size_t slab = static_cast<std::size_t>(ie) * timelevels + tl;
    const real* v = data.arrays.elem_state_v + slab * lev_pts * 2;
    const real* T = data.arrays.elem_state_T + slab * lev_pts;
    const real* dp = data.arrays.elem_state_dp3d + slab * lev_pts;
    for (int k = 0; k < nlev; ++k) {
      for (int f = 0; f < 4; ++f) out[f] << "[" << ie << ", " << k << "]\n";
      for (int i = 0; i < np; ++i) {
        for (int j = 0; j < np; ++j) {
          const std::size_t n = (static_cast<std::size_t>(k) * np + i) * np + j;
          out[0] << " " << v[2 * n];
          out[1] << " " << v[2 * n + 1];
          out[2] << " " << T[n];
          out[3] << " " << dp[n];
        }
        for (int f = 0; f < 4; ++f) out[f] << "\n";
      }
    }
  }
}

}  // namespace Homme
