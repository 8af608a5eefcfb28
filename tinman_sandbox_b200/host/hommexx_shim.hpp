// hommexx_shim.hpp — the HOMMEXX-style host interface of the reference's Kokkos variants on top of the C-ABI:
// the classes and member functions a Fortran-driven HOMME build calls to hand its arrays to the C++ side
// (compute_and_apply_rhs_test/cxx/level_vectorized_ppscan/ = "LV/"):
//
//   Homme::Control::init(nets, nete, num_elems, nm1, n0, np1, qn0, dt2, ps0, compute_diagnostics, eta_ave_w, hybrid_a)
//                                                                       LV/Control.hpp:13-17, LV/Control.cpp:5-29
//   Homme::Derivative::init(dvv)                                        LV/Derivative.hpp:15, LV/Derivative.cpp:11-23
//   Homme::Elements::init(num_elems)                                    LV/Elements.hpp:88
//   Homme::Elements::init_2d(D, Dinv, fcor, spheremp, metdet, phis)     LV/Elements.hpp:95-96, LV/Elements.cpp:48-99
//   Homme::Elements::pull_from_f90_pointers(state_v, state_t, state_dp3d, derived_phi, derived_pecnd,
//       derived_omega_p, derived_v, derived_eta_dot_dpdn, state_qdp)    LV/Elements.hpp:99-103, LV/Elements.cpp:154-292
//   Homme::Elements::push_to_f90_pointers(...same order...)             LV/Elements.hpp:111-115, LV/Elements.cpp:294-435
//   Homme::Elements::pull_3d / pull_4d / pull_eta_dot / pull_qdp and the push_* counterparts
//                                                                       LV/Elements.hpp:104-117
//   Homme::caar(control, elements, derivative)                          = Kokkos::parallel_for(policy, CaarFunctor(...)),
//                                                                         LV/kokkos_init.cpp:105-131
//
// Every pointer is an F90 flat pointer: the Fortran array as it lies in memory, element index slowest
// (CAAR_LAYOUT_F90 in include/caar_b200.h). The state lives on the GPU between pull and push, as in HOMMEXX.
// Same names, same argument order and meaning; no Kokkos. Header-only; link with libcaar_b200.so.
#ifndef CAAR_HOMMEXX_SHIM_HPP
#define CAAR_HOMMEXX_SHIM_HPP

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "caar_b200.h"

namespace Homme {

using Real = double;
using CRCPtr = const Real* const;
using F90Ptr = Real* const;
using CF90Ptr = const Real* const;

struct PhysicalConstants {  // LV/PhysicalConstants.hpp:9-17
  static constexpr Real Rwater_vapor = 461.5;
  static constexpr Real Rgas = 287.04;
  static constexpr Real cp = 1005.0;
  static constexpr Real kappa = Rgas / cp;
  static constexpr Real rrearth = 1.0 / 6.376e6;
};

// NUM_PHYSICAL_LEV is a compile-time constant of the reference (PLEV, LV/config.h.in:8, LV/Dimensions.hpp:36) and a
// run-time one here: a process-wide value, PLEV's default unless the caller (or Elements::init) sets another.
#ifndef PLEV
#define CAAR_SHIM_DEFAULT_PLEV 72
#else
#define CAAR_SHIM_DEFAULT_PLEV PLEV
#endif
inline int& num_physical_lev() {
  static int n = CAAR_SHIM_DEFAULT_PLEV;
  return n;
}

namespace detail {
[[noreturn]] inline void die(const char* what, int rc) {
  std::fprintf(stderr, "hommexx shim: %s failed (code %d): %s\n", what, rc, caar_last_error());
  std::abort();  // the reference's own failure convention
}
inline void ok(const char* what, int rc) {
  if (rc) die(what, rc);
}
}  // namespace detail

struct Control {
  int nets = 0, nete = 0, num_elems = 0, n0 = 0, nm1 = 0, np1 = 0, qn0 = -1, qsize = 0, compute_diagonstics = 0;
  int rsplit = 1;  // LV/Control.hpp:47-49: > 0 vertically Lagrangian
  Real dt = 0, eta_ave_w = 0, ps0 = 0;
  std::vector<Real> hybrid_a;
  // the reference's twelve arguments, in its order (LV/Control.hpp:13-17); hybrid_a has NUM_LEV_P =
  // num_physical_lev() + 1 entries (LV/Control.cpp:21-26)
  void init(const int nets_in, const int nete_in, const int num_elems_in, const int nm1_in, const int n0_in,
            const int np1_in, const int qn0_in, const Real dt_in, const Real ps0_in, const bool compute_diagonstics_in,
            const Real eta_ave_w_in, CRCPtr hybrid_a_ptr) {
    nets = nets_in; nete = nete_in; num_elems = num_elems_in; n0 = n0_in; nm1 = nm1_in; np1 = np1_in; qn0 = qn0_in;
    dt = dt_in; ps0 = ps0_in; compute_diagonstics = compute_diagonstics_in; eta_ave_w = eta_ave_w_in;
    hybrid_a.assign(hybrid_a_ptr, hybrid_a_ptr + num_physical_lev() + 1);
  }
};

class Derivative {
 public:
  void init(CF90Ptr& dvv_ptr) {  // deriv%Dvv(np,np) as stored by Fortran
    for (int i = 0; i < 16; ++i) m_dvv_f90[i] = dvv_ptr[i];
  }
  const Real* dvv_f90() const { return m_dvv_f90; }

 private:
  Real m_dvv_f90[16] = {};
};

class Elements {
 public:
  Elements() = default;
  Elements(const Elements&) = delete;
  ~Elements() {
    if (m_h) caar_destroy(m_h);
  }
  // nlev, qsize_d, timelevels are compile-time in the reference (LV/config.h.in) and run-time here
  void init(const int num_elems, const int nlev = num_physical_lev(), const int qsize_d = 1, const int timelevels = 3,
            const int device = 0) {
    num_physical_lev() = nlev;
    m_dims = caar_dims{num_elems, nlev, CAAR_NP, qsize_d, timelevels};
    detail::ok("caar_create", caar_create(&m_h, &m_dims, device));
  }
  int num_elems() const { return m_dims.nelem; }
  caar_handle handle() const { return m_h; }
  const caar_dims& dims() const { return m_dims; }

  void init_2d(CF90Ptr& D, CF90Ptr& Dinv, CF90Ptr& fcor, CF90Ptr& spheremp, CF90Ptr& metdet, CF90Ptr& phis) {
    caar_arrays a = {};
    a.elem_D = const_cast<Real*>(D); a.elem_Dinv = const_cast<Real*>(Dinv); a.elem_fcor = const_cast<Real*>(fcor);
    a.elem_spheremp = const_cast<Real*>(spheremp); a.elem_metdet = const_cast<Real*>(metdet);
    a.elem_state_phis = const_cast<Real*>(phis);
    a.elem_rmetdet = nullptr;  // HOMMEXX passes none: 1/metdet is formed on the device
    const unsigned mask = CAAR_F_D | CAAR_F_DINV | CAAR_F_FCOR | CAAR_F_SPHEREMP | CAAR_F_METDET | CAAR_F_RMETDET | CAAR_F_PHIS;
    detail::ok("caar_upload_layout", caar_upload_layout(m_h, &a, mask, CAAR_LAYOUT_F90));
  }

  void pull_from_f90_pointers(CF90Ptr& state_v, CF90Ptr& state_t, CF90Ptr& state_dp3d, CF90Ptr& derived_phi,
                              CF90Ptr& derived_pecnd, CF90Ptr& derived_omega_p, CF90Ptr& derived_v,
                              CF90Ptr& derived_eta_dot_dpdn, CF90Ptr& state_qdp) {
    caar_arrays a = view(const_cast<Real*>(state_v), const_cast<Real*>(state_t), const_cast<Real*>(state_dp3d),
                         const_cast<Real*>(derived_phi), const_cast<Real*>(derived_pecnd),
                         const_cast<Real*>(derived_omega_p), const_cast<Real*>(derived_v),
                         const_cast<Real*>(derived_eta_dot_dpdn), const_cast<Real*>(state_qdp));
    detail::ok("caar_upload_layout", caar_upload_layout(m_h, &a, k3d4d, CAAR_LAYOUT_F90));
  }

  // the four partial pulls / pushes the combined calls are made of (LV/Elements.hpp:104-117, LV/Elements.cpp:163-292,
  // 303-435): same names, same argument order
  void pull_3d(CF90Ptr& derived_phi, CF90Ptr& derived_pecnd, CF90Ptr& derived_omega_p, CF90Ptr& derived_v) {
    caar_arrays a = view(nullptr, nullptr, nullptr, const_cast<Real*>(derived_phi), const_cast<Real*>(derived_pecnd),
                         const_cast<Real*>(derived_omega_p), const_cast<Real*>(derived_v), nullptr, nullptr);
    detail::ok("caar_upload_layout", caar_upload_layout(m_h, &a, k3d, CAAR_LAYOUT_F90));
  }
  void pull_4d(CF90Ptr& state_v, CF90Ptr& state_t, CF90Ptr& state_dp3d) {
    caar_arrays a = view(const_cast<Real*>(state_v), const_cast<Real*>(state_t), const_cast<Real*>(state_dp3d), nullptr,
                         nullptr, nullptr, nullptr, nullptr, nullptr);
    detail::ok("caar_upload_layout", caar_upload_layout(m_h, &a, k4d, CAAR_LAYOUT_F90));
  }
  void pull_eta_dot(CF90Ptr& derived_eta_dot_dpdn) {
    caar_arrays a = view(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                         const_cast<Real*>(derived_eta_dot_dpdn), nullptr);
    detail::ok("caar_upload_layout", caar_upload_layout(m_h, &a, CAAR_F_ETA_DOT_DPDN, CAAR_LAYOUT_F90));
  }
  void pull_qdp(CF90Ptr& state_qdp) {
    caar_arrays a = view(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, const_cast<Real*>(state_qdp));
    detail::ok("caar_upload_layout", caar_upload_layout(m_h, &a, CAAR_F_QDP, CAAR_LAYOUT_F90));
  }
  void push_3d(F90Ptr& derived_phi, F90Ptr& derived_pecnd, F90Ptr& derived_omega_p, F90Ptr& derived_v) const {
    caar_arrays a = view(nullptr, nullptr, nullptr, derived_phi, derived_pecnd, derived_omega_p, derived_v, nullptr, nullptr);
    detail::ok("caar_download_layout", caar_download_layout(m_h, &a, k3d, CAAR_LAYOUT_F90));
  }
  void push_4d(F90Ptr& state_v, F90Ptr& state_t, F90Ptr& state_dp3d) const {
    caar_arrays a = view(state_v, state_t, state_dp3d, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    detail::ok("caar_download_layout", caar_download_layout(m_h, &a, k4d, CAAR_LAYOUT_F90));
  }
  void push_eta_dot(F90Ptr& derived_eta_dot_dpdn) const {
    caar_arrays a = view(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, derived_eta_dot_dpdn, nullptr);
    detail::ok("caar_download_layout", caar_download_layout(m_h, &a, CAAR_F_ETA_DOT_DPDN, CAAR_LAYOUT_F90));
  }
  void push_qdp(F90Ptr& state_qdp) const {
    caar_arrays a = view(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, state_qdp);
    detail::ok("caar_download_layout", caar_download_layout(m_h, &a, CAAR_F_QDP, CAAR_LAYOUT_F90));
  }

  void push_to_f90_pointers(F90Ptr& state_v, F90Ptr& state_t, F90Ptr& state_dp, F90Ptr& derived_phi, F90Ptr& derived_pecnd,
                            F90Ptr& derived_omega_p, F90Ptr& derived_v, F90Ptr& derived_eta_dot_dpdn,
                            F90Ptr& state_qdp) const {
    caar_arrays a = view(state_v, state_t, state_dp, derived_phi, derived_pecnd, derived_omega_p, derived_v,
                         derived_eta_dot_dpdn, state_qdp);
    detail::ok("caar_download_layout", caar_download_layout(m_h, &a, k3d4d, CAAR_LAYOUT_F90));
  }

 private:
  static constexpr unsigned k3d = CAAR_F_PHI | CAAR_F_PECND | CAAR_F_OMEGA_P | CAAR_F_VN0;
  static constexpr unsigned k4d = CAAR_F_V | CAAR_F_T | CAAR_F_DP3D;
  static constexpr unsigned k3d4d = CAAR_F_V | CAAR_F_T | CAAR_F_DP3D | CAAR_F_PHI | CAAR_F_PECND | CAAR_F_OMEGA_P |
                                    CAAR_F_VN0 | CAAR_F_ETA_DOT_DPDN | CAAR_F_QDP;
  static caar_arrays view(Real* v, Real* t, Real* dp, Real* phi, Real* pecnd, Real* omega_p, Real* dv, Real* eta, Real* qdp) {
    caar_arrays a = {};
    a.elem_state_v = v; a.elem_state_T = t; a.elem_state_dp3d = dp; a.elem_derived_phi = phi; a.elem_derived_pecnd = pecnd;
    a.elem_derived_omega_p = omega_p; a.elem_derived_vn0 = dv; a.elem_derived_eta_dot_dpdn = eta; a.elem_state_Qdp = qdp;
    return a;
  }
  caar_handle m_h = nullptr;
  caar_dims m_dims = {};
};

// One evaluation of compute_and_apply_rhs on the resident state: what Kokkos::parallel_for(policy, CaarFunctor(data,
// elements, deriv)) does in LV/kokkos_init.cpp:105-131. hybrid_a[0]*ps0 is the only use of hyai (PO:82).
inline void caar(const Control& data, Elements& elements, const Derivative& deriv, const int mode = CAAR_MODE_FAST) {
  const caar_constants c = {PhysicalConstants::rrearth, data.eta_ave_w, PhysicalConstants::cp,
                            PhysicalConstants::Rwater_vapor, PhysicalConstants::Rgas, PhysicalConstants::kappa};
  detail::ok("caar_set_params_f90",
             caar_set_params_f90(elements.handle(), &c, deriv.dvv_f90(), data.ps0, data.hybrid_a.data()));
  caar_control ctl = {data.nets, data.nete, data.n0, data.np1, data.nm1, data.qn0, data.dt};
  detail::ok("caar_run", caar_run(elements.handle(), &ctl, 1, mode));
  detail::ok("caar_sync", caar_sync(elements.handle()));
}

}  // namespace Homme
#endif
