// hommexx_shim_test.cpp — drives the HOMMEXX-style interface (hommexx_shim.hpp) the way a Fortran HOMME build would:
// every array is handed over as an F90 flat pointer (Fortran memory order, element index slowest), the state is
// pulled to the GPU, compute_and_apply_rhs runs `num_exec` times, the state is pushed back. The data are the
// reference's closed-form TestData, so the printed norms must be the reference driver's
// (tests/golden/pointers_only_stdout.txt; tests/test_host_gpu.py checks it).
//   hommexx_shim_test [num_elems=10] [num_exec=1] [fast|strict]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "closed_form_init.hpp"
#include "hommexx_shim.hpp"

namespace {

// C++ (pointers_only) order -> Fortran memory order and back, one element block at a time.
// kind 0: [i][j] <-> [j][i];  kind 1: [i][j][c] <-> [c][j][i];  kind 2: [i][j][a][b] <-> [b][a][j][i]
void relayout(const std::vector<double>& src, std::vector<double>& dst, int kind, bool to_f90) {
  const int B = 16 << kind;
  dst.resize(src.size());
  for (size_t blk = 0; blk < src.size() / B; ++blk)
    for (int q = 0; q < B; ++q) {
      int f;
      if (kind == 0) f = (q & 3) * 4 + (q >> 2);
      else if (kind == 1) f = (q & 1) * 16 + ((q >> 1) & 3) * 4 + (q >> 3);
      else f = ((q & 1) * 2 + ((q >> 1) & 1)) * 16 + ((q >> 2) & 3) * 4 + (q >> 4);
      if (to_f90) dst[blk * B + f] = src[blk * B + q];
      else dst[blk * B + q] = src[blk * B + f];
    }
}
int kind_of(int field) { return (field == 0 || field == 1) ? 2 : (field == 7 || field == 15) ? 1 : 0; }

}  // namespace

int main(int argc, char** argv) {
  const int E = argc > 1 ? std::atoi(argv[1]) : 10;
  const int nexec = argc > 2 ? std::atoi(argv[2]) : 1;
  const int mode = (argc > 3 && std::strcmp(argv[3], "strict") == 0) ? CAAR_MODE_STRICT : CAAR_MODE_FAST;
  const int L = 72;
  caar_host::HostData d;
  caar_host::init_data(d, E, L);
  enum { D, DINV, FCOR, MP, MET, RMET, DP, V, T, PHIS, QDP, ETA, OM, PHI, PEC, VN0 };
  std::vector<double> f[CAAR_NUM_FIELDS];  // the Fortran-side arrays (qsize_d = 1: Qdp block order is unchanged)
  for (int i = 0; i < CAAR_NUM_FIELDS; ++i) relayout(d.f[i], f[i], kind_of(i), true);
  double dvv_f90[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) dvv_f90[j * 4 + i] = d.dvv[i * 4 + j];

  Homme::Elements elements;
  elements.init(E, L);  // also sets Homme::num_physical_lev() (PLEV is compile-time in the reference)
  Homme::Control control;  // the reference's twelve arguments (LV/Control.hpp:13-17)
  control.init(d.ctl.nets, d.ctl.nete, E, d.ctl.nm1, d.ctl.n0, d.ctl.np1, d.ctl.qn0, d.ctl.dt2, d.ps0, false,
               d.c.eta_ave_w, d.hyai.data());
  Homme::Derivative deriv;
  deriv.init(dvv_f90);
  elements.init_2d(f[D].data(), f[DINV].data(), f[FCOR].data(), f[MP].data(), f[MET].data(), f[PHIS].data());
  if (argc > 4 && std::strcmp(argv[4], "partial") == 0) {  // the four partial pulls instead of the combined one
    elements.pull_3d(f[PHI].data(), f[PEC].data(), f[OM].data(), f[VN0].data());
    elements.pull_4d(f[V].data(), f[T].data(), f[DP].data());
    elements.pull_eta_dot(f[ETA].data());
    elements.pull_qdp(f[QDP].data());
  } else {
    elements.pull_from_f90_pointers(f[V].data(), f[T].data(), f[DP].data(), f[PHI].data(), f[PEC].data(), f[OM].data(),
                                    f[VN0].data(), f[ETA].data(), f[QDP].data());
  }
  for (int it = 0; it < nexec; ++it) Homme::caar(control, elements, deriv, mode);
  if (argc > 4 && std::strcmp(argv[4], "partial") == 0) {
    elements.push_3d(f[PHI].data(), f[PEC].data(), f[OM].data(), f[VN0].data());
    elements.push_4d(f[V].data(), f[T].data(), f[DP].data());
    elements.push_eta_dot(f[ETA].data());
    elements.push_qdp(f[QDP].data());
  } else {
    elements.push_to_f90_pointers(f[V].data(), f[T].data(), f[DP].data(), f[PHI].data(), f[PEC].data(), f[OM].data(),
                                  f[VN0].data(), f[ETA].data(), f[QDP].data());
  }

  // norms of v, T, dp3d at np1 straight from the Fortran-order arrays (a 2-norm does not care about the order)
  const size_t lev_pts = (size_t)L * 16;
  double acc[3] = {0, 0, 0};
  for (int ie = 0; ie < E; ++ie) {
    const size_t slab = (size_t)ie * 3 + d.ctl.np1;
    for (size_t n = 0; n < lev_pts * 2; ++n) acc[0] += f[V][slab * lev_pts * 2 + n] * f[V][slab * lev_pts * 2 + n];
    for (size_t n = 0; n < lev_pts; ++n) {
      acc[1] += f[T][slab * lev_pts + n] * f[T][slab * lev_pts + n];
      acc[2] += f[DP][slab * lev_pts + n] * f[DP][slab * lev_pts + n];
    }
  }
  std::printf("   ---> Norms:\n          ||v||_2  = %.17g\n          ||T||_2  = %.17g\n          ||dp||_2 = %.17g\n",
              std::sqrt(acc[0]), std::sqrt(acc[1]), std::sqrt(acc[2]));
  // and the round trip of the layout: element 0, level 0 of T(np1) in Fortran order is the transpose of the C++ order
  std::vector<double> back;
  relayout(f[T], back, 0, false);
  std::printf("T(np1)[0][0][1][2] (C++ order) = %.17g == Fortran T(2,3,1,np1) = %.17g\n",
              back[((size_t)d.ctl.np1 * L) * 16 + 1 * 4 + 2], f[T][((size_t)d.ctl.np1 * L) * 16 + 2 * 4 + 1]);
  return 0;
}
