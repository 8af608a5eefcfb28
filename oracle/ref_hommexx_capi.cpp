// oracle/ref_hommexx_capi.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's two HOMMEXX prototypes, compiled UNMODIFIED from /root/reference against the
// serial Kokkos stand-in oracle/kokkos_stub/Kokkos_Core.hpp (Kokkos is not in this image):
//
//   -DHX_TV=0  compute_and_apply_rhs_test/cxx/level_vectorized_ppscan  ("LV", fields [ie][..][igp][jgp][lev-pack])
//   -DHX_TV=1  compute_and_apply_rhs_test/cxx/tiled_vectorized_ppscan  ("TV", fields [ie][..][lev-pack][igp][jgp])
//
// together with that variant's own Control.cpp, Derivative.cpp and Elements.cpp (oracle/Makefile, targets lv / tv).
// This file holds NO arithmetic of the path: it moves numbers between flat arrays and the reference's Views and
// calls the reference's functions:
//
//   hx_caar_f90        Elements::init / init_2d / pull_from_f90_pointers, Control::init, Derivative::init,
//                      CaarFunctor::operator() per element, Elements::push_to_f90_pointers
//                      (LV/Elements.cpp:8-99,154-435, LV/Control.cpp:5-28, LV/Derivative.cpp:11-23,
//                       LV/CaarFunctor.hpp:549-563) — the F90 flat-pointer boundary of SURVEY §8f rank 1, executed by
//                      the reference's own code
//   hx_sphere_op       gradient_sphere, divergence_sphere, vorticity_sphere_vector, divergence_sphere_wk, laplace_simple,
//                      laplace_tensor, laplace_tensor_replace, divergence_sphere_update
//                      (LV/SphereOperators.hpp:227-636, TV/SphereOperators.hpp:221-549)
//   hx_preq_vertadv    CaarFunctor::preq_vertadv (LV/CaarFunctor.hpp:504-547, TV/CaarFunctor.hpp:499-542)
//   hx_euler_step      EulerStepFunctor::operator() (TV/EulerStepFunctor.hpp:33-70) — TV only: the LV copy of that
//                      functor indexes its [NP][NP][NUM_LEV] views in TV order and assigns them to
//                      [NUM_LEV][NP][NP]-typed views, which Kokkos' (and this stub's) static extent check rejects at
//                      compile time; nothing in the reference's build ever instantiates it (LV/CMakeLists.txt:21-28).
//
// Flat-array convention of hx_sphere_op / hx_preq_vertadv / hx_euler_step ("hx order"): the index order of the TV views,
// [ie][...][lev][comp][igp][jgp] with HOMMEXX's (igp, jgp), independent of the variant compiled. The Python side
// (oracle/harness.py) maps the pointers_only arrays to it.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "CaarFunctor.hpp"
#if HX_TV
#include "EulerStepFunctor.hpp"
#endif

// the reference's GPTL timers (profiling.hpp:15-18) are not part of the path: link-time no-ops
extern "C" int GPTLstart(const char*) { return 0; }
extern "C" int GPTLstop(const char*) { return 0; }

using namespace Homme;

namespace {

inline Real& lane(Scalar& s, int iv) { return s[iv]; }

// level-field accessors that hide the two view layouts
#if HX_TV
#define HX3(view, lev, i, j) (view)((lev) / VECTOR_SIZE, (i), (j))[(lev) % VECTOR_SIZE]
#define HX4(view, lev, c, i, j) (view)((lev) / VECTOR_SIZE, (c), (i), (j))[(lev) % VECTOR_SIZE]
using Scalar3 = Scalar[NUM_LEV][NP][NP];
using Scalar4 = Scalar[NUM_LEV][2][NP][NP];
using Scalar3P = Scalar[NUM_LEV_P][NP][NP];
#else
#define HX3(view, lev, i, j) (view)((i), (j), (lev) / VECTOR_SIZE)[(lev) % VECTOR_SIZE]
#define HX4(view, lev, c, i, j) (view)((c), (i), (j), (lev) / VECTOR_SIZE)[(lev) % VECTOR_SIZE]
using Scalar3 = Scalar[NP][NP][NUM_LEV];
using Scalar4 = Scalar[2][NP][NP][NUM_LEV];
#endif

void load3(ExecViewManaged<Scalar3> v, const double* hx) {
  for (int lev = 0; lev < NUM_PHYSICAL_LEV; ++lev)
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) HX3(v, lev, i, j) = hx[(lev * NP + i) * NP + j];
}
void store3(ExecViewManaged<Scalar3> v, double* hx) {
  for (int lev = 0; lev < NUM_PHYSICAL_LEV; ++lev)
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) hx[(lev * NP + i) * NP + j] = HX3(v, lev, i, j);
}
void load4(ExecViewManaged<Scalar4> v, const double* hx) {
  for (int lev = 0; lev < NUM_PHYSICAL_LEV; ++lev)
    for (int c = 0; c < 2; ++c)
      for (int i = 0; i < NP; ++i)
        for (int j = 0; j < NP; ++j) HX4(v, lev, c, i, j) = hx[((lev * 2 + c) * NP + i) * NP + j];
}
void store4(ExecViewManaged<Scalar4> v, double* hx) {
  for (int lev = 0; lev < NUM_PHYSICAL_LEV; ++lev)
    for (int c = 0; c < 2; ++c)
      for (int i = 0; i < NP; ++i)
        for (int j = 0; j < NP; ++j) hx[((lev * 2 + c) * NP + i) * NP + j] = HX4(v, lev, c, i, j);
}

}  // namespace

extern "C" {

int hx_variant() { return HX_TV; }
int hx_nlev() { return NUM_PHYSICAL_LEV; }
int hx_qsize_d() { return QSIZE_D; }
int hx_vector_size() { return VECTOR_SIZE; }

// One or more calls of the HOMMEXX compute_and_apply_rhs on arrays in Fortran memory order, in place.
void hx_caar_f90(int nelem, const double* D, const double* Dinv, const double* fcor, const double* spheremp,
                 const double* metdet, const double* phis, double* v, double* T, double* dp3d, double* phi, double* pecnd,
                 double* omega_p, double* vn0, double* eta_dot_dpdn, double* qdp, int nets, int nete, int nm1, int n0,
                 int np1, int qn0, double dt2, double ps0, double eta_ave_w, const double* hyai, const double* dvv_f90,
                 int ncalls) {
  Elements elements;
  elements.init(nelem);
  elements.init_2d(D, Dinv, fcor, spheremp, metdet, phis);
  elements.pull_from_f90_pointers(v, T, dp3d, phi, pecnd, omega_p, vn0, eta_dot_dpdn, qdp);
  Control control;
  control.init(nets, nete, nelem, nm1, n0, np1, qn0, dt2, ps0, false, eta_ave_w, hyai);
  Derivative deriv;
  deriv.init(dvv_f90);
  CaarFunctor func(control, elements, deriv);
  for (int c = 0; c < ncalls; ++c)
    for (int ie = nets; ie < nete; ++ie) {
      TeamMember team(ie, nelem);
      func(team);
    }
  elements.push_to_f90_pointers(v, T, dp3d, phi, pecnd, omega_p, vn0, eta_dot_dpdn, qdp);
}

// Sphere operators on one element (geometry of element 0 of a one-element Elements), every level.
//   op 0 gradient_sphere          in  [L][4][4]      out [L][2][4][4]
//   op 1 divergence_sphere        in  [L][2][4][4]   out [L][4][4]
//   op 2 vorticity_sphere_vector  in  [L][2][4][4]   out [L][4][4]
//   op 3 divergence_sphere_wk     in  [L][2][4][4]   out [L][4][4]
//   op 4 laplace_simple           in  [L][4][4]      out [L][4][4]
//   op 5 laplace_tensor           in  [L][4][4]      out [L][4][4]
//   op 6 laplace_tensor_replace   in  [L][4][4]      out [L][4][4]   (the reference's in-place form)
//   op 7 divergence_sphere_update in  [L][2][4][4]   out [L][4][4] = beta*out + alpha*div(in)   (out is in/out)
// d, dinv, tensorvisc: [2][2][4][4] in the reference's view order; metdet, spheremp: [4][4]; dvv: [4][4] view order.
int hx_sphere_op(int op, const double* d, const double* dinv, const double* metdet, const double* spheremp,
                 const double* tensorvisc, const double* dvv, const double* in, double* out, double alpha, double beta) {
  Elements elements;
  elements.init(1);
  ExecViewManaged<Real * [2][2][NP][NP]> tvisc("tensorVisc", 1);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int i = 0; i < NP; ++i)
        for (int j = 0; j < NP; ++j) {
          const int q = ((a * 2 + b) * NP + i) * NP + j;
          elements.m_d(0, a, b, i, j) = d[q];
          elements.m_dinv(0, a, b, i, j) = dinv[q];
          tvisc(0, a, b, i, j) = tensorvisc ? tensorvisc[q] : 0.0;
        }
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      elements.m_metdet(0, i, j) = metdet[i * NP + j];
      elements.m_spheremp(0, i, j) = spheremp[i * NP + j];
    }
  ExecViewManaged<Real[NP][NP]> dvv_view("dvv");
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) dvv_view(i, j) = dvv[i * NP + j];

  ExecViewManaged<Scalar3> s_in("scalar in"), s_out("scalar out");
  ExecViewManaged<Scalar4> v_in("vector in"), v_out("vector out / gradient temporary");
  const bool vector_in = (op == 1 || op == 2 || op == 3 || op == 7);
  if (vector_in) load4(v_in, in); else load3(s_in, in);
  if (op == 7) load3(s_out, out);
  if (op == 6) load3(s_out, in);

  TeamMember team(0, 1);
  KernelVariables kv(team);
#if HX_TV
  for (int ilev = 0; ilev < NUM_LEV; ++ilev) {
    kv.ilev = ilev;
    switch (op) {
      case 0: gradient_sphere(kv, elements.m_dinv, dvv_view, s_in, v_out); break;
      case 1: divergence_sphere(kv, elements.m_dinv, elements.m_metdet, dvv_view, v_in, s_out); break;
      case 2: vorticity_sphere_vector(kv, elements.m_d, elements.m_metdet, dvv_view, v_in, s_out); break;
      case 3: divergence_sphere_wk(kv, elements.m_dinv, elements.m_spheremp, dvv_view, v_in, s_out); break;
      case 4: laplace_simple(kv, elements.m_dinv, elements.m_spheremp, dvv_view, v_out, s_in, s_out); break;
      case 5: laplace_tensor(kv, elements.m_dinv, elements.m_spheremp, dvv_view, tvisc, v_out, s_in, s_out); break;
      case 6: laplace_tensor_replace(kv, elements.m_dinv, elements.m_spheremp, dvv_view, tvisc, v_out, s_out); break;
      case 7:
        divergence_sphere_update(kv, alpha, beta, Homme::subview(elements.m_dinv, 0), Homme::subview(elements.m_metdet, 0),
                                 dvv_view, v_in, s_out);
        break;
      default: return 1;
    }
  }
#else
  ExecViewManaged<Scalar * [2][NP][NP][NUM_LEV]> buf("sphere_buf", 1);
  switch (op) {
    case 0: gradient_sphere(kv, elements.m_dinv, dvv_view, s_in, buf, v_out); break;
    case 1: divergence_sphere(kv, elements.m_dinv, elements.m_metdet, dvv_view, v_in, buf, s_out); break;
    case 2: vorticity_sphere_vector(kv, elements.m_d, elements.m_metdet, dvv_view, v_in, buf, s_out); break;
    case 3: divergence_sphere_wk(kv, elements.m_dinv, elements.m_spheremp, dvv_view, v_in, buf, s_out); break;
    case 4: laplace_simple(kv, elements.m_dinv, elements.m_spheremp, dvv_view, v_out, s_in, buf, s_out); break;
    case 5: laplace_tensor(kv, elements.m_dinv, elements.m_spheremp, dvv_view, tvisc, v_out, s_in, buf, s_out); break;
    case 6: laplace_tensor_replace(kv, elements.m_dinv, elements.m_spheremp, dvv_view, tvisc, v_out, buf, s_out); break;
    case 7:
      divergence_sphere_update(kv, alpha, beta, Homme::subview(elements.m_dinv, 0), Homme::subview(elements.m_metdet, 0),
                               dvv_view, v_in, buf, s_out);
      break;
    default: return 1;
  }
#endif
  if (op == 0) store4(v_out, out); else store3(s_out, out);
  return 0;
}

// CaarFunctor::preq_vertadv: its arguments are [lev][..][j][i]-typed views in BOTH variants.
void hx_preq_vertadv(const double* T, const double* v, const double* eta_dp_deta, const double* rpdel, double* T_vadv,
                     double* v_vadv) {
  static_assert(VECTOR_SIZE == 1, "hx_preq_vertadv wraps the flat arrays directly: scalar packs only");
  ExecViewManaged<Scalar[NUM_LEV][NP][NP]> vT("T"), vr("rpdel"), vTa("T_vadv");
  ExecViewManaged<Scalar[NUM_LEV][2][NP][NP]> vv("v"), vva("v_vadv");
  ExecViewManaged<Scalar[NUM_LEV_P][NP][NP]> ve("eta_dp_deta");
  for (int k = 0; k < NUM_LEV; ++k)
    for (int q = 0; q < NP * NP; ++q) {
      vT(k, q / NP, q % NP)[0] = T[k * 16 + q];
      vr(k, q / NP, q % NP)[0] = rpdel[k * 16 + q];
      for (int h = 0; h < 2; ++h) vv(k, h, q / NP, q % NP)[0] = v[(k * 2 + h) * 16 + q];
    }
  for (int k = 0; k < NUM_LEV_P; ++k)
    for (int q = 0; q < NP * NP; ++q) ve(k, q / NP, q % NP)[0] = eta_dp_deta[k * 16 + q];
  Control control;
  Elements elements;
  Derivative deriv;
  CaarFunctor func(control, elements, deriv);
  TeamMember team(0, 1);
  func.preq_vertadv(team, vT, vv, ve, vr, vTa, vva);
  for (int k = 0; k < NUM_LEV; ++k)
    for (int q = 0; q < NP * NP; ++q) {
      T_vadv[k * 16 + q] = vTa(k, q / NP, q % NP)[0];
      for (int h = 0; h < 2; ++h) v_vadv[(k * 2 + h) * 16 + q] = vva(k, h, q / NP, q % NP)[0];
    }
}

#if HX_TV
// EulerStepFunctor::operator() for every element. Arrays in view order:
//   dinv [E][2][2][4][4], metdet [E][4][4], dvv [4][4], vstar [E][L][2][4][4], qdp [E][2][QSIZE_D][L][4][4],
//   qtens (out) [E][QSIZE_D][L][4][4]
void hx_euler_step(int nelem, int qsize, int qn0, double dt, const double* dinv, const double* metdet, const double* dvv,
                   const double* vstar, const double* qdp, double* qtens) {
  static_assert(VECTOR_SIZE == 1, "hx_euler_step fills the views level by level: scalar packs only");
  Elements& elements = get_elements();  // the functor takes the singletons (TV/EulerStepFunctor.hpp:18-24)
  elements.init(nelem);
  for (int ie = 0; ie < nelem; ++ie) {
    for (int q = 0; q < 64; ++q) elements.m_dinv(ie, q / 32, (q / 16) % 2, (q / 4) % 4, q % 4) = dinv[ie * 64 + q];
    for (int q = 0; q < 16; ++q) elements.m_metdet(ie, q / 4, q % 4) = metdet[ie * 16 + q];
    for (int k = 0; k < NUM_LEV; ++k)
      for (int h = 0; h < 2; ++h)
        for (int q = 0; q < 16; ++q)
          elements.buffers.vstar(ie, k, h, q / 4, q % 4)[0] = vstar[((size_t(ie) * NUM_LEV + k) * 2 + h) * 16 + q];
    for (int t = 0; t < Q_NUM_TIME_LEVELS; ++t)
      for (int iq = 0; iq < QSIZE_D; ++iq)
        for (int k = 0; k < NUM_LEV; ++k)
          for (int q = 0; q < 16; ++q)
            elements.m_qdp(ie, t, iq, k, q / 4, q % 4)[0] =
                qdp[(((size_t(ie) * Q_NUM_TIME_LEVELS + t) * QSIZE_D + iq) * NUM_LEV + k) * 16 + q];
  }
  Derivative& deriv = get_derivative();
  {
    double f90[16];  // Derivative::init reads Fortran memory order and fills dvv(igp,jgp) = ptr[igp*4+jgp]
    for (int q = 0; q < 16; ++q) f90[q] = dvv[q];
    deriv.init(f90);
  }
  Control control;
  control.nets = 0;
  control.nete = nelem;
  control.num_elems = nelem;
  control.qn0 = qn0;
  control.qsize = qsize;
  control.dt = dt;
  EulerStepFunctor func(control);
  for (int ie = 0; ie < nelem; ++ie) {
    TeamMember team(ie, nelem);
    func(team);
  }
  for (int ie = 0; ie < nelem; ++ie)
    for (int iq = 0; iq < QSIZE_D; ++iq)
      for (int k = 0; k < NUM_LEV; ++k)
        for (int q = 0; q < 16; ++q)
          qtens[((size_t(ie) * QSIZE_D + iq) * NUM_LEV + k) * 16 + q] = elements.buffers.qtens(ie, iq, k, q / 4, q % 4)[0];
}
#endif

}  // extern "C"
