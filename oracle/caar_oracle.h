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

/* the three printed norms (PO/compute_and_apply_rhs.cpp:354-399) of time level tl */
void caar_oracle_norms(int nlev, int ntl, double* const* arrays, int nets, int nete, int tl,
                       double out3[3]);

/* x = a*x + b*y (saxpby_test/cxx/common.cpp:3-15), nthreads pthreads; returns wall seconds */
double caar_oracle_saxpby(double a, double b, double* x, const double* y, size_t n, int sweeps,
                          int nthreads);

#ifdef __cplusplus
}
#endif
#endif
