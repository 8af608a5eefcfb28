/* oracle/caar_oracle.h — TEST INFRASTRUCTURE ONLY. Interface of the CPU restatement (caar_oracle.c).
 * Mirrors oracle/ref_capi.cpp so the tests can swap the real reference (_ref) and the restatement.
 * Arrays are passed as a table of 16 pointers in the order of struct Arrays
 * (reference: compute_and_apply_rhs_test/cxx/pointers_only/data_structures.hpp:18-44). */
#ifndef CAAR_ORACLE_H
#define CAAR_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

size_t caar_oracle_field_count(int field, int nelem, int nlev, int qsize_d, int ntl);

/* closed-form synthetic init (PO/data_structures.cpp:38-92,117-163) */
void caar_oracle_init(int nelem, int nlev, int qsize_d, int ntl, double* const* arrays, int* ctl6i,
                      double* dt2, double* consts6, double* dvv16, double* ps0, double* hyai);

/* ncalls evaluations of compute_and_apply_rhs (PO/compute_and_apply_rhs.cpp:15-278) on [nets,nete),
 * split over nthreads pthreads by contiguous element ranges. Returns wall seconds. */
double caar_oracle_run(int nlev, int qsize_d, int ntl, double* const* arrays, const int* ctl6i, double dt2,
                       const double* consts6, const double* dvv16, double ps0, const double* hyai,
                       int ncalls, int nthreads);

/* the same with the Eulerian vertical coordinate (rsplit == 0: eta_dot_dpdn from the divergence sum and hybi,
 * vertical advection of T and v; fortran/routine_extracted.F90:227-262,325-334,515-517 and preq_vertadv,
 * level_vectorized_ppscan/CaarFunctor.hpp:504-547). hybi has nlev+1 entries. PARITY UNPINNED: the C++ reference
 * does not implement this branch and the Fortran one cannot be compiled here. */
double caar_oracle_run_eulerian(int nlev, int qsize_d, int ntl, double* const* arrays, const int* ctl6i, double dt2,
                                const double* consts6, const double* dvv16, double ps0, const double* hyai,
                                const double* hybi, int ncalls, int nthreads);

/* the three printed norms (PO/compute_and_apply_rhs.cpp:354-399) of time level tl */
void caar_oracle_norms(int nlev, int ntl, double* const* arrays, int nets, int nete, int tl,
                       double out3[3]);

/* one 4x4 level of the reference's divergence_sphere (PO/sphere_operators.cpp:50-89): v [4][4][2], dinv [4][4][2][2] */
void caar_oracle_divergence_sphere(const double* v, const double* dvv16, const double* dinv, const double* metdet,
                                   const double* rmetdet, double rrearth, double* div);

/* tracer step after CAAR: qtens = Qdp(qn0) - dt*divergence_sphere(vstar*Qdp(qn0)) for tracers [0,qsize), every
 * level, elements [nets,nete) (level_vectorized_ppscan/EulerStepFunctor.hpp:33-66).
 * vstar [E][L][4][4][2], qtens [E][qsize_d][L][4][4] */
void caar_oracle_euler_step(int nlev, int qsize_d, double* const* arrays, const double* vstar, double* qtens,
                            int nets, int nete, int qn0, int qsize, double dt, const double* dvv16,
                            double rrearth);

/* preq_vertadv (level_vectorized_ppscan/CaarFunctor.hpp:504-547): T, rpdel, T_vadv [L][4][4]; v, v_vadv [L][4][4][2];
 * eta_dp_deta [L+1][4][4]. The Eulerian branch of caar_oracle_run_eulerian calls exactly this. */
void caar_oracle_preq_vertadv(int nlev, const double* T, const double* v, const double* eta_dp_deta, const double* rpdel,
                              double* T_vadv, double* v_vadv);

/* weak-form operators behind hyperviscosity (level_vectorized_ppscan/SphereOperators.hpp:493-636) in the pointers_only
 * conventions: op 0 divergence_sphere_wk (vin [E][L][4][4][2]), op 1 laplace_simple, op 2 laplace_tensor (sin
 * [E][L][4][4], tensorvisc [E][4][4][2][2]); out [E][L][4][4] */
void caar_oracle_sphere_wk(int op, int nlev, double* const* arrays, const double* vin, const double* sin,
                           const double* tensorvisc, double* out, int nets, int nete, const double* dvv16,
                           double rrearth);
void caar_oracle_gradient_sphere(const double* s, const double* dvv16, const double* dinv, double rrearth, double* ds);
void caar_oracle_vorticity_sphere(const double* v, const double* dvv16, const double* d, const double* rmetdet,
                                  double rrearth, double* vort);

/* x = a*x + b*y (saxpby_test/cxx/common.cpp:3-15), nthreads pthreads; returns wall seconds */
double caar_oracle_saxpby(double a, double b, double* x, const double* y, size_t n, int sweeps,
                          int nthreads);

#ifdef __cplusplus
}
#endif
#endif
