/* oracle/caar_oracle.c — TEST INFRASTRUCTURE ONLY. NOT part of the product, never linked into
 * libcaar_b200.so, never called by the product path (only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it).
 *
 * A plain-C restatement ("port") of the reference's compute_and_apply_rhs hot path with run-time
 * nlev, written to round EXACTLY like the reference's `g++ -std=c++11 -O3` x86-64 build: same
 * operation order, no FMA contraction (built with -ffp-contract=off), IEEE division.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this file
 *   (1) bit-for-bit against the real reference compiled from /root/reference (oracle/_ref, see
 *       oracle/Makefile) on every mutated array, for PLEV=72 and PLEV=128,
 *   (2) against the reference's own golden vectors Ttest/v1test/v2test
 *       (compute_and_apply_rhs_test/fortran/test_mod.F90:8,299,594; committed as tests/golden/*.npy),
 *   (3) against the norms the reference driver prints (tests/golden/pointers_only_stdout.txt).
 *
 * Reference map ("PO/" = compute_and_apply_rhs_test/cxx/pointers_only/):
 *   grad_sphere()        PO/sphere_operators.cpp:9-48
 *   div_sphere()         PO/sphere_operators.cpp:50-89
 *   vort_sphere()        PO/sphere_operators.cpp:91-129
 *   hydrostatic()        PO/compute_and_apply_rhs.cpp:280-312   (preq_hydrostatic)
 *   omega_ps()           PO/compute_and_apply_rhs.cpp:314-352   (preq_omega_ps)
 *   rhs_element()        PO/compute_and_apply_rhs.cpp:74-258    (body of the element loop)
 *   kahan_norm()         PO/compute_and_apply_rhs.cpp:354-370   (compute_norm)
 *   caar_oracle_norms()  PO/compute_and_apply_rhs.cpp:372-399   (print_results_2norm)
 *   caar_oracle_init()   PO/data_structures.cpp:38-92,117-163   (init_data of all five structs)
 */
#include "caar_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { NP = 4, PTS = 16 };

enum {
  F_D, F_DINV, F_FCOR, F_SPHEREMP, F_METDET, F_RMETDET, F_DP3D, F_V, F_T, F_PHIS, F_QDP,
  F_ETA_DOT_DPDN, F_OMEGA_P, F_PHI, F_PECND, F_VN0
};

typedef struct {
  int nlev, qsize_d, ntl;
  double* const* a;          /* 16 array pointers */
  int nets, nete, n0, np1, nm1, qn0;
  double dt2;
  double rrearth, eta_ave_w, cp, Rwv, Rgas, kappa;
  const double* dvv;         /* [4][4] row-major */
  double ps0;
  const double* hyai;
  /* vertical coordinate: rsplit > 0 = vertically Lagrangian (the only branch the C++ reference runs),
   * rsplit == 0 = Eulerian (fortran/routine_extracted.F90:227-262), which needs hybi[nlev+1] */
  int rsplit;
  const double* hybi;
} ctx_t;

size_t caar_oracle_field_count(int f, int E, int L, int Q, int ntl) {
  const size_t e = (size_t)E;
  switch (f) {
    case F_D: case F_DINV: return e * PTS * 4;
    case F_FCOR: case F_SPHEREMP: case F_METDET: case F_RMETDET: case F_PHIS: return e * PTS;
    case F_DP3D: case F_T: return e * ntl * L * PTS;
    case F_V: return e * ntl * L * PTS * 2;
    case F_QDP: return e * Q * 2 * L * PTS;
    case F_ETA_DOT_DPDN: return e * (L + 1) * PTS;
    case F_OMEGA_P: case F_PHI: case F_PECND: return e * L * PTS;
    case F_VN0: return e * L * PTS * 2;
  }
  return 0;
}

/* ---- sphere operators on one 4x4 level of one element ------------------------------------ */

/* ds[i][j][c] = sum over the covariant derivatives, PO/sphere_operators.cpp:21-47 */
static void grad_sphere(const double* s, const double* dvv, const double* dinv, double rrearth,
                        double* ds) {
  double a[NP][NP], b[NP][NP];
  for (int j = 0; j < NP; ++j)
    for (int l = 0; l < NP; ++l) {
      double sx = 0, sy = 0;
      for (int i = 0; i < NP; ++i) {
        sx += dvv[i * NP + l] * s[i * NP + j];
        sy += dvv[i * NP + l] * s[j * NP + i];
      }
      a[l][j] = sx * rrearth;
      b[j][l] = sy * rrearth;
    }
  for (int j = 0; j < NP; ++j)
    for (int i = 0; i < NP; ++i) {
      const double* di = dinv + (i * NP + j) * 4; /* [2][2] */
      ds[(i * NP + j) * 2 + 0] = di[0] * a[i][j] + di[2] * b[i][j];
      ds[(i * NP + j) * 2 + 1] = di[1] * a[i][j] + di[3] * b[i][j];
    }
}

/* PO/sphere_operators.cpp:62-88 */
static void div_sphere(const double* v, const double* dvv, const double* dinv, const double* metdet,
                       const double* rmetdet, double rrearth, double* div) {
  double g[NP][NP][2];
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      const double* di = dinv + (i * NP + j) * 4;
      const double v0 = v[(i * NP + j) * 2], v1 = v[(i * NP + j) * 2 + 1];
      g[i][j][0] = metdet[i * NP + j] * (di[0] * v0 + di[1] * v1);
      g[i][j][1] = metdet[i * NP + j] * (di[2] * v0 + di[3] * v1);
    }
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      double dudx = 0., dvdy = 0.;
      for (int k = 0; k < NP; ++k) {
        dudx += dvv[k * NP + i] * g[k][j][0];
        dvdy += dvv[k * NP + j] * g[i][k][1];
      }
      div[i * NP + j] = (dudx + dvdy) * rmetdet[i * NP + j] * rrearth;
    }
}

/* PO/sphere_operators.cpp:102-128 */
static void vort_sphere(const double* v, const double* dvv, const double* d, const double* rmetdet,
                        double rrearth, double* vort) {
  double c[NP][NP][2];
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      const double* dd = d + (i * NP + j) * 4;
      const double v0 = v[(i * NP + j) * 2], v1 = v[(i * NP + j) * 2 + 1];
      c[i][j][0] = dd[0] * v0 + dd[2] * v1;
      c[i][j][1] = dd[1] * v0 + dd[3] * v1;
    }
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      double dudy = 0., dvdx = 0.;
      for (int k = 0; k < NP; ++k) {
        dvdx += dvv[k * NP + i] * c[k][j][1];
        dudy += dvv[k * NP + j] * c[i][k][0];
      }
      vort[i * NP + j] = (dvdx - dudy) * rmetdet[i * NP + j] * rrearth;
    }
}

/* divergence_sphere_wk, the weak (integrated-by-parts) divergence behind hyperviscosity:
 * level_vectorized_ppscan/SphereOperators.hpp:493-535 (= tiled_vectorized_ppscan/SphereOperators.hpp:421-455), in the
 * pointers_only index convention (HOMMEXX view (igp,jgp) == pointers_only [jgp][igp], tensors (x,y) == [.][.][y][x]):
 *   vtemp[i][j][c] = Dinv[i][j][c][0]*v0 + Dinv[i][j][c][1]*v1
 *   div[m][n] = - sum_j ( spheremp[j][n]*vtemp[j][n][0]*Dvv[m][j] + spheremp[m][j]*vtemp[m][j][1]*Dvv[n][j] ) * rrearth
 * accumulated from 0 by repeated `dd -= term` in j order, every product left to right, as the reference writes it
 * (its accumulator `Scalar dd;` is zero-initialised by the vendored Vector's default constructor,
 * level_vectorized_ppscan/vector/KokkosKernels_Vector_SIMD.hpp:32-37). Pinned to the reference's own code run under
 * oracle/kokkos_stub (tests/test_oracle.py). */
static void div_sphere_wk(const double* v, const double* dvv, const double* dinv, const double* spheremp,
                          double rrearth, double* div) {
  double t[NP][NP][2];
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) {
      const double* di = dinv + (i * NP + j) * 4;
      const double v0 = v[(i * NP + j) * 2], v1 = v[(i * NP + j) * 2 + 1];
      t[i][j][0] = di[0] * v0 + di[1] * v1;
      t[i][j][1] = di[2] * v0 + di[3] * v1;
    }
  for (int m = 0; m < NP; ++m)
    for (int n = 0; n < NP; ++n) {
      double dd = 0.0;
      for (int j = 0; j < NP; ++j)
        dd -= (spheremp[j * NP + n] * t[j][n][0] * dvv[m * NP + j] + spheremp[m * NP + j] * t[m][j][1] * dvv[n * NP + j]) *
              rrearth;
      div[m * NP + n] = dd;
    }
}

/* laplace_simple / laplace_tensor / laplace_tensor_replace (level_vectorized_ppscan/SphereOperators.hpp:537-636):
 * gradient_sphere, optionally g <- tensorVisc . g (gv[c] = tv[i][j][c][0]*g0 + tv[i][j][c][1]*g1), then
 * divergence_sphere_wk. tensorvisc == NULL selects laplace_simple. */
static void laplace_wk(const double* s, const double* dvv, const double* dinv, const double* spheremp,
                       const double* tensorvisc, double rrearth, double* lap) {
  double g[PTS * 2];
  grad_sphere(s, dvv, dinv, rrearth, g);
  if (tensorvisc)
    for (int q = 0; q < PTS; ++q) {
      const double* tv = tensorvisc + q * 4;
      const double g0 = g[q * 2], g1 = g[q * 2 + 1];
      g[q * 2] = tv[0] * g0 + tv[1] * g1;
      g[q * 2 + 1] = tv[2] * g0 + tv[3] * g1;
    }
  div_sphere_wk(g, dvv, dinv, spheremp, rrearth, lap);
}

/* preq_vertadv, the vertical advection of T and v (CCM2 3.b.1): level_vectorized_ppscan/CaarFunctor.hpp:504-547.
 * T, rpdel, T_vadv [L][16]; v, v_vadv [L][16][2]; eta_dp_deta [L+1][16]. Pinned to the reference's own code. */
static void preq_vertadv(int L, const double* T, const double* v, const double* eta, const double* rpdel,
                         double* T_vadv, double* v_vadv) {
  for (int k = 0; k < L; ++k)
    for (int q = 0; q < PTS; ++q) {
      const size_t n = (size_t)k * PTS + q;
      const double facp = 0.5 * rpdel[n] * eta[n + PTS];
      const double facm = 0.5 * rpdel[n] * eta[n];
      if (k == 0) {
        T_vadv[n] = facp * (T[n + PTS] - T[n]);
        for (int h = 0; h < 2; ++h) v_vadv[n * 2 + h] = facp * (v[(n + PTS) * 2 + h] - v[n * 2 + h]);
      } else if (k < L - 1) {
        T_vadv[n] = facp * (T[n + PTS] - T[n]) + facm * (T[n] - T[n - PTS]);
        for (int h = 0; h < 2; ++h)
          v_vadv[n * 2 + h] = facp * (v[(n + PTS) * 2 + h] - v[n * 2 + h]) + facm * (v[n * 2 + h] - v[(n - PTS) * 2 + h]);
      } else {
        T_vadv[n] = facm * (T[n] - T[n - PTS]);
        for (int h = 0; h < 2; ++h) v_vadv[n * 2 + h] = facm * (v[n * 2 + h] - v[(n - PTS) * 2 + h]);
      }
    }
}

/* ---- vertical integrals -------------------------------------------------------------------- */

/* reverse (bottom-up) sum; phii is an [L][16] scratch. PO/compute_and_apply_rhs.cpp:287-311 */
static void hydrostatic(int L, const double* phis, const double* T_v, const double* p, const double* dp,
                        double Rgas, double* phii, double* phi) {
  for (int q = 0; q < PTS; ++q) {
    int k = L - 1;
    double hkk = 0.5 * dp[k * PTS + q] / p[k * PTS + q];
    double hkl = 2.0 * hkk;
    phii[k * PTS + q] = Rgas * T_v[k * PTS + q] * hkl;
    phi[k * PTS + q] = phis[q] + Rgas * T_v[k * PTS + q] * hkk;
    for (k = L - 2; k > 0; --k) {
      hkk = 0.5 * dp[k * PTS + q] / p[k * PTS + q];
      hkl = 2.0 * hkk;
      phii[k * PTS + q] = phii[(k + 1) * PTS + q] + Rgas * T_v[k * PTS + q] * hkl;
      phi[k * PTS + q] = phis[q] + phii[(k + 1) * PTS + q] + Rgas * T_v[k * PTS + q] * hkk;
    }
    hkk = 0.5 * dp[q] / p[q];
    phi[q] = phis[q] + phii[PTS + q] + Rgas * T_v[q] * hkk;
  }
}

/* forward (top-down) running sum of divdp. PO/compute_and_apply_rhs.cpp:319-351 */
static void omega_ps(int L, const double* p, const double* vgrad_p, const double* divdp, double* omega) {
  for (int q = 0; q < PTS; ++q) {
    double ckk = 0.5 / p[q];
    double term = divdp[q];
    omega[q] = vgrad_p[q] / p[q] - ckk * term;
    double suml = term;
    for (int k = 1; k < L - 1; ++k) {
      ckk = 0.5 / p[k * PTS + q];
      const double ckl = 2.0 * ckk;
      term = divdp[k * PTS + q];
      omega[k * PTS + q] = vgrad_p[k * PTS + q] / p[k * PTS + q] - ckl * suml - ckk * term;
      suml += term;
    }
    const int k = L - 1;
    ckk = 0.5 / p[k * PTS + q];
    const double ckl = 2.0 * ckk;
    term = divdp[k * PTS + q];
    omega[k * PTS + q] = vgrad_p[k * PTS + q] / p[k * PTS + q] - ckl * suml - ckk * term;
  }
}

/* ---- one element ---------------------------------------------------------------------------- */

typedef struct {
  double *p, *grad_p, *vgrad_p, *vdp, *divdp, *vort, *T_v, *omega, *phii, *vt1, *vt2, *tt;
  double* eta; /* [L+1][16] interface values of eta_dot_dpdn (Eulerian branch) */
  double *rpdel, *T_vadv, *v_vadv; /* Eulerian branch: 1/dp and the output of preq_vertadv */
} scratch_t;

static int scratch_alloc(scratch_t* w, int L) {
  const size_t n = (size_t)L * PTS;
  double* base = (double*)calloc(n * 14 + n + PTS + 4 * n, sizeof(double));
  if (!base) return -1;
  w->p = base;            w->grad_p = base + n;      w->vgrad_p = base + 3 * n;
  w->vdp = base + 4 * n;  w->divdp = base + 6 * n;   w->vort = base + 7 * n;
  w->T_v = base + 8 * n;  w->omega = base + 9 * n;   w->phii = base + 10 * n;
  w->vt1 = base + 11 * n; w->vt2 = base + 12 * n;    w->tt = base + 13 * n;
  w->eta = base + 14 * n;
  w->rpdel = base + 15 * n + PTS; w->T_vadv = w->rpdel + n; w->v_vadv = w->T_vadv + n;
  return 0;
}

static void rhs_element(const ctx_t* c, int ie, scratch_t* w) {
  const int L = c->nlev;
  const size_t lf = (size_t)L * PTS;                 /* one scalar level-field */
  double* const* A = c->a;
  const double* D = A[F_D] + (size_t)ie * PTS * 4;
  const double* Dinv = A[F_DINV] + (size_t)ie * PTS * 4;
  const double* fcor = A[F_FCOR] + (size_t)ie * PTS;
  const double* spheremp = A[F_SPHEREMP] + (size_t)ie * PTS;
  const double* metdet = A[F_METDET] + (size_t)ie * PTS;
  const double* rmetdet = A[F_RMETDET] + (size_t)ie * PTS;
  const double* phis = A[F_PHIS] + (size_t)ie * PTS;
  const double* dp_n0 = A[F_DP3D] + ((size_t)ie * c->ntl + c->n0) * lf;
  const double* v_n0 = A[F_V] + ((size_t)ie * c->ntl + c->n0) * lf * 2;
  const double* T_n0 = A[F_T] + ((size_t)ie * c->ntl + c->n0) * lf;
  double* vn0 = A[F_VN0] + (size_t)ie * lf * 2;
  double* phi = A[F_PHI] + (size_t)ie * lf;
  double* omega_p = A[F_OMEGA_P] + (size_t)ie * lf;
  double* eta_dot = A[F_ETA_DOT_DPDN] + (size_t)ie * (L + 1) * PTS;
  const double* pecnd = A[F_PECND] + (size_t)ie * lf;

  /* A: mid-level pressure, top-down (PO:76-97) */
  for (int q = 0; q < PTS; ++q) w->p[q] = c->hyai[0] * c->ps0 + 0.5 * dp_n0[q];
  for (int k = 1; k < L; ++k)
    for (int q = 0; q < PTS; ++q)
      w->p[k * PTS + q] = w->p[(k - 1) * PTS + q] + 0.5 * dp_n0[(k - 1) * PTS + q] + 0.5 * dp_n0[k * PTS + q];

  /* B: level-local horizontal operators (PO:101-124) */
  for (int k = 0; k < L; ++k) {
    double* gp = w->grad_p + (size_t)k * PTS * 2;
    grad_sphere(w->p + k * PTS, c->dvv, Dinv, c->rrearth, gp);
    for (int q = 0; q < PTS; ++q) {
      const double v1 = v_n0[(k * PTS + q) * 2], v2 = v_n0[(k * PTS + q) * 2 + 1];
      w->vgrad_p[k * PTS + q] = v1 * gp[q * 2] + v2 * gp[q * 2 + 1];
      w->vdp[(k * PTS + q) * 2] = v1 * dp_n0[k * PTS + q];
      w->vdp[(k * PTS + q) * 2 + 1] = v2 * dp_n0[k * PTS + q];
      vn0[(k * PTS + q) * 2] += c->eta_ave_w * w->vdp[(k * PTS + q) * 2];
      vn0[(k * PTS + q) * 2 + 1] += c->eta_ave_w * w->vdp[(k * PTS + q) * 2 + 1];
    }
    div_sphere(w->vdp + (size_t)k * PTS * 2, c->dvv, Dinv, metdet, rmetdet, c->rrearth, w->divdp + k * PTS);
    vort_sphere(v_n0 + (size_t)k * PTS * 2, c->dvv, D, rmetdet, c->rrearth, w->vort + k * PTS);
  }

  /* C: virtual temperature (PO:126-156). kappa_star is the constant kappa everywhere. */
  if (c->qn0 == -1) {
    for (size_t n = 0; n < lf; ++n) w->T_v[n] = T_n0[n];
  } else {
    const double* Qdp = A[F_QDP] + (((size_t)ie * c->qsize_d + 0) * 2 + c->qn0) * lf;
    for (size_t n = 0; n < lf; ++n) {
      const double Qt = Qdp[n] / dp_n0[n];
      w->T_v[n] = T_n0[n] * (1.0 + (c->Rwv / c->Rgas - 1.0) * Qt);
    }
  }

  /* D, E: the two vertical integrals (PO:161-162) */
  hydrostatic(L, phis, w->T_v, w->p, dp_n0, c->Rgas, w->phii, phi);
  omega_ps(L, w->p, w->vgrad_p, w->divdp, w->omega);

  /* Eulerian branch (rsplit == 0), which the C++ reference leaves out ("-0" placeholders, PO:222-231) and the
   * Fortran reference specifies: eta_dot_dpdn at the L+1 interfaces from the running sum of divdp and hybi,
   * fortran/routine_extracted.F90:233-254. NOT pinned by any runnable reference (no Fortran compiler, no known
   * answers): "parity unpinned" for this branch. */
  if (c->rsplit == 0) {
    double sdot[PTS];
    for (int q = 0; q < PTS; ++q) sdot[q] = 0.0;
    for (int k = 0; k < L; ++k)
      for (int q = 0; q < PTS; ++q) {
        sdot[q] += w->divdp[k * PTS + q];
        w->eta[(k + 1) * PTS + q] = sdot[q];
      }
    for (int k = 0; k < L - 1; ++k)
      for (int q = 0; q < PTS; ++q)
        w->eta[(k + 1) * PTS + q] = c->hybi[k + 1] * sdot[q] - w->eta[(k + 1) * PTS + q];
    for (int q = 0; q < PTS; ++q) {
      w->eta[q] = 0.0;
      w->eta[(size_t)L * PTS + q] = 0.0;
    }
  }

  /* F: accumulate the derived fields (PO:164-183). eta_dot_dpdn_tmp is identically zero on the Lagrangian
   * branch; fortran/routine_extracted.F90:270-277 on the Eulerian one. */
  if (c->rsplit == 0) {
    for (size_t n = 0; n < lf + PTS; ++n) eta_dot[n] += c->eta_ave_w * w->eta[n];
    for (size_t n = 0; n < lf; ++n) omega_p[n] += c->eta_ave_w * w->omega[n];
  } else {
    const double zero = 0.0;
    for (size_t n = 0; n < lf; ++n) {
      eta_dot[n] += c->eta_ave_w * zero;
      omega_p[n] += c->eta_ave_w * w->omega[n];
    }
    for (int q = 0; q < PTS; ++q) eta_dot[lf + q] += c->eta_ave_w * zero;
  }

  if (c->rsplit == 0) { /* rpdel = 1/dp (fortran/routine_extracted.F90:120) */
    for (size_t n = 0; n < lf; ++n) w->rpdel[n] = 1.0 / dp_n0[n];
    preq_vertadv(L, T_n0, v_n0, w->eta, w->rpdel, w->T_vadv, w->v_vadv);
  }

  /* G: tendencies (PO:187-234). v_vadv and T_vadv are identically zero in the reference. */
  for (int k = 0; k < L; ++k) {
    double Ephi[PTS], gT[PTS * 2], gE[PTS * 2], vgrad_T[PTS];
    for (int q = 0; q < PTS; ++q) {
      const double v1 = v_n0[(k * PTS + q) * 2], v2 = v_n0[(k * PTS + q) * 2 + 1];
      Ephi[q] = 0.5 * (v1 * v1 + v2 * v2) + phi[k * PTS + q] + pecnd[k * PTS + q];
    }
    grad_sphere(T_n0 + k * PTS, c->dvv, Dinv, c->rrearth, gT);
    for (int q = 0; q < PTS; ++q) {
      const double v1 = v_n0[(k * PTS + q) * 2], v2 = v_n0[(k * PTS + q) * 2 + 1];
      vgrad_T[q] = v1 * gT[q * 2] + v2 * gT[q * 2 + 1];
    }
    grad_sphere(Ephi, c->dvv, Dinv, c->rrearth, gE);
    const double* gp = w->grad_p + (size_t)k * PTS * 2;
    if (c->rsplit == 0) {
      /* vertical advection from preq_vertadv (computed for the whole column before this loop), then the tendencies
       * with the Fortran signs (fortran/routine_extracted.F90:325-334: ttens = -T_vadv - vgrad_T + kappa*T_v*omega_p) */
      for (int q = 0; q < PTS; ++q) {
        const size_t n = (size_t)k * PTS + q;
        const double T_vadv = w->T_vadv[n], v_vadv0 = w->v_vadv[n * 2], v_vadv1 = w->v_vadv[n * 2 + 1];
        const double gpterm = w->T_v[n] / w->p[n];
        const double glnps1 = c->Rgas * gpterm * gp[q * 2];
        const double glnps2 = c->Rgas * gpterm * gp[q * 2 + 1];
        const double v1 = v_n0[n * 2], v2 = v_n0[n * 2 + 1];
        w->vt1[n] = -v_vadv0 + v2 * (fcor[q] + w->vort[n]) - gE[q * 2] - glnps1;
        w->vt2[n] = -v_vadv1 - v1 * (fcor[q] + w->vort[n]) - gE[q * 2 + 1] - glnps2;
        w->tt[n] = -T_vadv - vgrad_T[q] + c->kappa * w->T_v[n] * w->omega[n];
      }
      continue;
    }
    for (int q = 0; q < PTS; ++q) {
      const double v_vadv0 = 0.0, v_vadv1 = 0.0, T_vadv = 0.0;
      const double gpterm = w->T_v[k * PTS + q] / w->p[k * PTS + q];
      const double glnps1 = c->Rgas * gpterm * gp[q * 2];
      const double glnps2 = c->Rgas * gpterm * gp[q * 2 + 1];
      const double v1 = v_n0[(k * PTS + q) * 2], v2 = v_n0[(k * PTS + q) * 2 + 1];
      w->vt1[k * PTS + q] = -v_vadv0 + v2 * (fcor[q] + w->vort[k * PTS + q]) - gE[q * 2] - glnps1;
      w->vt2[k * PTS + q] = -v_vadv1 - v1 * (fcor[q] + w->vort[k * PTS + q]) - gE[q * 2 + 1] - glnps2;
      w->tt[k * PTS + q] = T_vadv - vgrad_T[q] + c->kappa * w->T_v[k * PTS + q] * w->omega[k * PTS + q];
    }
  }

  /* H: apply (PO:236-257) */
  {
    double* v_np1 = A[F_V] + ((size_t)ie * c->ntl + c->np1) * lf * 2;
    double* T_np1 = A[F_T] + ((size_t)ie * c->ntl + c->np1) * lf;
    double* dp_np1 = A[F_DP3D] + ((size_t)ie * c->ntl + c->np1) * lf;
    const double* v_nm1 = A[F_V] + ((size_t)ie * c->ntl + c->nm1) * lf * 2;
    const double* T_nm1 = A[F_T] + ((size_t)ie * c->ntl + c->nm1) * lf;
    const double* dp_nm1 = A[F_DP3D] + ((size_t)ie * c->ntl + c->nm1) * lf;
    for (int k = 0; k < L; ++k)
      for (int q = 0; q < PTS; ++q) {
        const size_t n = (size_t)k * PTS + q;
        v_np1[n * 2] = spheremp[q] * (v_nm1[n * 2] + c->dt2 * w->vt1[n]);
        v_np1[n * 2 + 1] = spheremp[q] * (v_nm1[n * 2 + 1] + c->dt2 * w->vt2[n]);
        T_np1[n] = spheremp[q] * (T_nm1[n] + c->dt2 * w->tt[n]);
        if (c->rsplit == 0) /* fortran/routine_extracted.F90:515-517 */
          dp_np1[n] = spheremp[q] * (dp_nm1[n] - c->dt2 * (w->divdp[n] + w->eta[n + PTS] - w->eta[n]));
        else
          dp_np1[n] = spheremp[q] * (dp_nm1[n] - c->dt2 * w->divdp[n]);
      }
  }
}

/* ---- driver ----------------------------------------------------------------------------------- */

typedef struct {
  ctx_t c;
  int ncalls;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  scratch_t w;
  if (scratch_alloc(&w, j->c.nlev)) return (void*)1;
  for (int n = 0; n < j->ncalls; ++n)
    for (int ie = j->c.nets; ie < j->c.nete; ++ie) rhs_element(&j->c, ie, &w);
  free(w.p);
  return NULL;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double run_impl(int nlev, int qsize_d, int ntl, double* const* arrays, const int* ctl, double dt2,
                       const double* k6, const double* dvv16, double ps0, const double* hyai, int rsplit,
                       const double* hybi, int ncalls, int nthreads);

double caar_oracle_run(int nlev, int qsize_d, int ntl, double* const* arrays, const int* ctl, double dt2,
                       const double* k6, const double* dvv16, double ps0, const double* hyai, int ncalls,
                       int nthreads) {
  return run_impl(nlev, qsize_d, ntl, arrays, ctl, dt2, k6, dvv16, ps0, hyai, 1, NULL, ncalls, nthreads);
}

/* the Eulerian (rsplit == 0) branch: fortran/routine_extracted.F90:227-262,325-334,515-517 */
double caar_oracle_run_eulerian(int nlev, int qsize_d, int ntl, double* const* arrays, const int* ctl, double dt2,
                                const double* k6, const double* dvv16, double ps0, const double* hyai,
                                const double* hybi, int ncalls, int nthreads) {
  return run_impl(nlev, qsize_d, ntl, arrays, ctl, dt2, k6, dvv16, ps0, hyai, 0, hybi, ncalls, nthreads);
}

static double run_impl(int nlev, int qsize_d, int ntl, double* const* arrays, const int* ctl, double dt2,
                       const double* k6, const double* dvv16, double ps0, const double* hyai, int rsplit,
                       const double* hybi, int ncalls, int nthreads) {
  ctx_t c;
  memset(&c, 0, sizeof c);
  c.rsplit = rsplit; c.hybi = hybi;
  c.nlev = nlev; c.qsize_d = qsize_d; c.ntl = ntl; c.a = arrays;
  c.nets = ctl[0]; c.nete = ctl[1]; c.n0 = ctl[2]; c.np1 = ctl[3]; c.nm1 = ctl[4]; c.qn0 = ctl[5];
  c.dt2 = dt2;
  c.rrearth = k6[0]; c.eta_ave_w = k6[1]; c.cp = k6[2]; c.Rwv = k6[3]; c.Rgas = k6[4]; c.kappa = k6[5];
  c.dvv = dvv16; c.ps0 = ps0; c.hyai = hyai;
  const int n = c.nete - c.nets;
  if (n <= 0 || nlev < 2) return 0.0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = n;
  job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].c = c;
    jobs[t].ncalls = ncalls;
    jobs[t].c.nets = c.nets + (int)((long long)n * t / nthreads);
    jobs[t].c.nete = c.nets + (int)((long long)n * (t + 1) / nthreads);
  }
  const double t0 = now_s();
  if (nthreads == 1) {
    worker(&jobs[0]);
  } else {
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  const double t1 = now_s();
  free(jobs);
  free(th);
  return t1 - t0;
}

/* Kahan-compensated sum of squares, then sqrt (PO:354-370) */
static double kahan_norm(const double* f, size_t n) {
  double norm = 0, comp = 0;
  for (size_t i = 0; i < n; ++i) {
    const double y = f[i] * f[i] - comp;
    const double t = norm + y;
    comp = (t - norm) - y;
    norm = t;
  }
  return sqrt(norm);
}

void caar_oracle_norms(int nlev, int ntl, double* const* arrays, int nets, int nete, int tl, double out[3]) {
  const size_t lf = (size_t)nlev * PTS;
  double vn = 0, tn = 0, dn = 0;
  for (int ie = nets; ie < nete; ++ie) {
    vn += pow(kahan_norm(arrays[F_V] + ((size_t)ie * ntl + tl) * lf * 2, lf * 2), 2);
    tn += pow(kahan_norm(arrays[F_T] + ((size_t)ie * ntl + tl) * lf, lf), 2);
    dn += pow(kahan_norm(arrays[F_DP3D] + ((size_t)ie * ntl + tl) * lf, lf), 2);
  }
  out[0] = sqrt(vn);
  out[1] = sqrt(tn);
  out[2] = sqrt(dn);
}

/* ---- tracer step after CAAR ---------------------------------------
 * qtens[ie][iq][k] = Qdp[ie][iq][qn0][k] - dt * divergence_sphere(vstar[ie][k] * Qdp[ie][iq][qn0][k])
 * Composition: level_vectorized_ppscan/EulerStepFunctor.hpp:33-66 (v_buf = vstar*qdp, q_buf = qdp, then
 * divergence_sphere_update(alpha = -dt, beta = 1): q_buf = beta*q_buf + alpha*div(v_buf),
 * level_vectorized_ppscan/SphereOperators.hpp:362-403). Operator: the reference's tested
 * divergence_sphere, PO/sphere_operators.cpp:50-89 (div_sphere above, pinned bit-exactly to oracle/_ref).
 * The reference never calls EulerStepFunctor from a driver and holds no known answers for it: the
 * composition is pinned only by this restatement ("parity unpinned" for the composition, pinned for the
 * operator). Layouts follow the pointers_only conventions: vstar [E][L][4][4][2] (as derived_vn0),
 * qtens [E][qsize_d][L][4][4]. */
void caar_oracle_divergence_sphere(const double* v, const double* dvv16, const double* dinv, const double* metdet,
                                   const double* rmetdet, double rrearth, double* div) {
  div_sphere(v, dvv16, dinv, metdet, rmetdet, rrearth, div);
}

void caar_oracle_euler_step(int nlev, int qsize_d, double* const* arrays, const double* vstar, double* qtens,
                            int nets, int nete, int qn0, int qsize, double dt, const double* dvv16,
                            double rrearth) {
  const int L = nlev;
  for (int ie = nets; ie < nete; ++ie) {
    const double* dinv = arrays[F_DINV] + (size_t)ie * PTS * 4;
    const double* metdet = arrays[F_METDET] + (size_t)ie * PTS;
    const double* rmetdet = arrays[F_RMETDET] + (size_t)ie * PTS;
    for (int iq = 0; iq < qsize; ++iq)
      for (int k = 0; k < L; ++k) {
        const double* q = arrays[F_QDP] + ((((size_t)ie * qsize_d + iq) * 2 + qn0) * L + k) * PTS;
        const double* vs = vstar + ((size_t)ie * L + k) * PTS * 2;
        double* out = qtens + (((size_t)ie * qsize_d + iq) * L + k) * PTS;
        double vq[PTS * 2], div[PTS];
        for (int n = 0; n < PTS; ++n) {
          vq[2 * n] = vs[2 * n] * q[n];
          vq[2 * n + 1] = vs[2 * n + 1] * q[n];
        }
        div_sphere(vq, dvv16, dinv, metdet, rmetdet, rrearth, div);
        for (int n = 0; n < PTS; ++n) {
          double t = q[n];
          t *= 1.0;
          t += (-dt) * div[n];
          out[n] = t;
        }
      }
  }
}


/* ---- the weak-form operators and preq_vertadv, exported for the tests and as the oracle of caar_sphere_wk ---- */
void caar_oracle_preq_vertadv(int nlev, const double* T, const double* v, const double* eta_dp_deta, const double* rpdel,
                              double* T_vadv, double* v_vadv) {
  preq_vertadv(nlev, T, v, eta_dp_deta, rpdel, T_vadv, v_vadv);
}

/* op 0: divergence_sphere_wk of vin [E][L][4][4][2]; op 1: laplace_simple, op 2: laplace_tensor of sin [E][L][4][4]
 * (tensorvisc [E][4][4][2][2]); out [E][L][4][4]; elements [nets,nete). Uses the Dinv / spheremp arrays of the table. */
void caar_oracle_sphere_wk(int op, int nlev, double* const* arrays, const double* vin, const double* sin,
                           const double* tensorvisc, double* out, int nets, int nete, const double* dvv16,
                           double rrearth) {
  for (int ie = nets; ie < nete; ++ie) {
    const double* dinv = arrays[F_DINV] + (size_t)ie * PTS * 4;
    const double* spheremp = arrays[F_SPHEREMP] + (size_t)ie * PTS;
    for (int k = 0; k < nlev; ++k) {
      const size_t n = ((size_t)ie * nlev + k) * PTS;
      if (op == 0) div_sphere_wk(vin + n * 2, dvv16, dinv, spheremp, rrearth, out + n);
      else laplace_wk(sin + n, dvv16, dinv, spheremp, op == 2 ? tensorvisc + (size_t)ie * PTS * 4 : NULL, rrearth, out + n);
    }
  }
}

/* one level of gradient_sphere / vorticity_sphere (PO/sphere_operators.cpp:9-48, 91-129), for the operator pins */
void caar_oracle_gradient_sphere(const double* s, const double* dvv16, const double* dinv, double rrearth, double* ds) {
  grad_sphere(s, dvv16, dinv, rrearth, ds);
}
void caar_oracle_vorticity_sphere(const double* v, const double* dvv16, const double* d, const double* rmetdet,
                                  double rrearth, double* vort) {
  vort_sphere(v, dvv16, d, rmetdet, rrearth, vort);
}

/* closed-form synthetic fields; 1-based index values as in the reference (PO/data_structures.cpp:38-92) */
void caar_oracle_init(int E, int L, int Q, int ntl, double* const* a, int* ctl, double* dt2, double* k6,
                      double* dvv16, double* ps0, double* hyai) {
  for (int f = 0; f < 16; ++f) memset(a[f], 0, caar_oracle_field_count(f, E, L, Q, ntl) * sizeof(double));
  for (int ie = 0; ie < E; ++ie)
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) {
        const double e1 = ie + 1, i1 = i + 1, j1 = j + 1;
        const size_t q = (size_t)ie * PTS + i * NP + j;
        a[F_FCOR][q] = sin(i1 + j1);
        a[F_METDET][q] = i1 * j1;
        a[F_RMETDET][q] = 1. / a[F_METDET][q];
        a[F_SPHEREMP][q] = 2 * i1;
        a[F_PHIS][q] = i1 + j1;
        a[F_D][q * 4 + 0] = 1.0; a[F_D][q * 4 + 3] = 2.0;
        a[F_DINV][q * 4 + 0] = 1.0; a[F_DINV][q * 4 + 3] = 0.5;
        for (int k = 0; k < L; ++k) {
          const double k1 = k + 1;
          const size_t n = ((size_t)ie * L + k) * PTS + i * NP + j;
          a[F_PHI][n] = cos(i1 + 3 * j1) + k1;
          a[F_VN0][n * 2] = 1.0;
          a[F_VN0][n * 2 + 1] = 1.0;
          a[F_PECND][n] = 1.0;
          a[F_OMEGA_P][n] = j1 * j1;
          a[F_QDP][(((size_t)ie * Q * 2) * L + k) * PTS + i * NP + j] = 1.0 + sin(i1 * j1 * k1);
          for (int t = 0; t < ntl; ++t) {
            const double t1 = t + 1;
            const size_t m = (((size_t)ie * ntl + t) * L + k) * PTS + i * NP + j;
            a[F_DP3D][m] = 10.0 * k1 + e1 + i1 + j1 + t1;
            a[F_V][m * 2] = 1.0 + 0.5 * k1 + i1 + j1 + 0.2 * e1 + 2.0 * t1;
            a[F_V][m * 2 + 1] = 1.0 + 0.5 * k1 + i1 + j1 + 0.2 * e1 + 3.0 * t1;
            a[F_T][m] = 1000.0 - k1 - i1 - j1 + 0.1 * e1 + t1;
          }
        }
      }
  ctl[0] = 0; ctl[1] = E; ctl[2] = 0; ctl[3] = 1; ctl[4] = 2; ctl[5] = 0;
  *dt2 = 1.0;
  k6[3] = 461.5;             /* Rwater_vapor */
  k6[4] = 287.04;            /* Rgas */
  k6[2] = 1005.0;            /* cp */
  k6[5] = k6[4] / k6[2];     /* kappa */
  k6[0] = 1.0 / 6.376e6;     /* rrearth */
  k6[1] = 1.0;               /* eta_ave_w */
  *ps0 = 10.0;
  for (int i = 0; i <= L; ++i) hyai[i] = L + 1 - i;
  /* GLL derivative matrix for np=4, stored transposed into Dvv (PO/data_structures.cpp:150-163) */
  static const double vals[16] = {-3.0000000000000000, -0.80901699437494745, 0.30901699437494745,
                                  -0.50000000000000000, 4.0450849718747373,  0.00000000000000000,
                                  -1.11803398874989490, 1.54508497187473700, -1.5450849718747370,
                                  1.11803398874989490,  0.00000000000000000, -4.04508497187473730,
                                  0.5000000000000000,   -0.30901699437494745, 0.80901699437494745,
                                  3.000000000000000000};
  for (int i = 0; i < NP; ++i)
    for (int j = 0; j < NP; ++j) dvv16[i * NP + j] = vals[j * NP + i];
}

/* ---- saxpby ----------------------------------------------------------------------------------- */
typedef struct { double a, b; double* x; const double* y; size_t lo, hi; int sweeps; } sax_t;
static void* sax_worker(void* arg) {
  sax_t* s = (sax_t*)arg;
  for (int it = 0; it < s->sweeps; ++it)
    for (size_t i = s->lo; i < s->hi; ++i) s->x[i] = s->a * s->x[i] + s->b * s->y[i];
  return NULL;
}
double caar_oracle_saxpby(double a, double b, double* x, const double* y, size_t n, int sweeps, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  sax_t* js = (sax_t*)calloc((size_t)nthreads, sizeof(sax_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
  const double t0 = now_s();
  for (int t = 0; t < nthreads; ++t) {
    js[t].a = a; js[t].b = b; js[t].x = x; js[t].y = y; js[t].sweeps = sweeps;
    js[t].lo = n * t / nthreads; js[t].hi = n * (t + 1) / nthreads;
    pthread_create(&th[t], NULL, sax_worker, &js[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  const double t1 = now_s();
  free(js); free(th);
  return t1 - t0;
}
