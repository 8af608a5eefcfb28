// caar_fused_ldg.cu — first-generation fused kernel (plain LDG/STG, no TMA), kept for nlev values without a
// TMA-pipelined instance (8, 16, 32, 64) and as an A/B baseline (CAAR_MODE_FAST_LDG): the whole of compute_and_apply_rhs for one element in ONE kernel,
// one HBM pass: every input is read once, every output written once, all 18 reference temporaries
// (PO/compute_and_apply_rhs.cpp:18-35) live in registers.
//
// Work decomposition ("R4"): one CTA per element, 4*nlev threads. Thread t owns level k = t/4 and GLL
// row igp = t%4 of that level, i.e. the 4 points jgp=0..3 — 32 contiguous bytes of every scalar
// level-field and 64 contiguous bytes of the interleaved (u,v) fields, so a warp reads 1 KB / 2 KB
// contiguous per field. A warp therefore holds 8 consecutive levels.
//
//  * sphere operators (PO/sphere_operators.cpp:9-129): the derivative along jgp is thread-local
//    (16 FMAs against Dvv from the constant bank); the derivative along igp needs the other three rows of
//    the level, which sit in lanes lane^1, lane^2, lane^3: 3 xor-shuffles per value.
//  * vertical integrals (pressure PO:76-97, preq_omega_ps PO:314-352, preq_hydrostatic PO:280-312):
//    warp-shuffle scans across the 8 levels of a warp (lane stride 4: offsets 4, 8, 16), then a carry
//    across the nlev/8 warps through 3 x (nlev/8) x 16 doubles of shared memory. Two __syncthreads
//    per element in total.
//  * divisions: one reciprocal of p serves hkk, ckk, vgrad_p/p and T_v/p (PO:300,333,336,219).
//
// Rounding differs from the reference (FMA contraction, tree-ordered sums, shared reciprocal): results
// agree to ~1e-14 relative, checked at 1e-12 per field in tests/test_parity_gpu.py.
#include "caar_device.cuh"

namespace caar {
namespace {

constexpr unsigned FULL = 0xffffffffu;

struct Row {  // the 4 points (jgp = 0..3) of one GLL row of one level
  double x[4];
};

__device__ __forceinline__ Row ld_row(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  Row r;
  r.x[0] = a.x; r.x[1] = a.y; r.x[2] = b.x; r.x[3] = b.y;
  return r;
}
__device__ __forceinline__ void st_row(double* p, const Row& r) {
  *reinterpret_cast<double2*>(p) = make_double2(r.x[0], r.x[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(r.x[2], r.x[3]);
}
// interleaved [jgp][2] -> two rows
__device__ __forceinline__ void ld_row2(const double* p, Row& u, Row& w) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double2 a = *reinterpret_cast<const double2*>(p + 2 * j);
    u.x[j] = a.x;
    w.x[j] = a.y;
  }
}
__device__ __forceinline__ void st_row2(double* p, const Row& u, const Row& w) {
#pragma unroll
  for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(p + 2 * j) = make_double2(u.x[j], w.x[j]);
}

// Derivative along igp: out[j] = sum_m Dvv[m][r] * s_m[j], rows m of this level live in lanes lane^x.
// cx[x] = Dvv[r^x][r].
__device__ __forceinline__ Row deriv_i(const Row& s, const double (&cx)[4]) {
  Row o;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double acc = cx[0] * s.x[j];
#pragma unroll
    for (int x = 1; x < 4; ++x) acc = fma(cx[x], __shfl_xor_sync(FULL, s.x[j], x), acc);
    o.x[j] = acc;
  }
  return o;
}
// Derivative along jgp: out[l] = sum_m Dvv[m][l] * s[m] (thread-local; Dvv from the constant bank)
__device__ __forceinline__ Row deriv_j(const Row& s, const double* __restrict__ dvv) {
  Row o;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    double acc = dvv[0 * 4 + l] * s.x[0];
#pragma unroll
    for (int m = 1; m < 4; ++m) acc = fma(dvv[m * 4 + l], s.x[m], acc);
    o.x[l] = acc;
  }
  return o;
}

struct Geo {       // per-thread geometry of its GLL row; Dinv pre-scaled by rrearth
  double di[4][4]; // [jgp][2*a+b] = Dinv[igp][jgp][a][b] * rrearth
};

// gradient_sphere (PO/sphere_operators.cpp:9-48) for this thread's row
__device__ __forceinline__ void gradient(const Row& s, const Geo& g, const double (&cx)[4],
                                         const double* __restrict__ dvv, Row& g0, Row& g1) {
  const Row a = deriv_i(s, cx);   // v1[igp][jgp]
  const Row b = deriv_j(s, dvv);  // v2[igp][jgp]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    g0.x[j] = fma(g.di[j][0], a.x[j], g.di[j][2] * b.x[j]);
    g1.x[j] = fma(g.di[j][1], a.x[j], g.di[j][3] * b.x[j]);
  }
}

// inclusive scan over the 8 levels of a warp (same igp => lane stride 4)
__device__ __forceinline__ double scan_down(double v, int lane) {  // towards larger k
#pragma unroll
  for (int d = 4; d < 32; d <<= 1) {
    const double t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}
__device__ __forceinline__ double scan_up(double v, int lane) {  // towards smaller k
#pragma unroll
  for (int d = 4; d < 32; d <<= 1) {
    const double t = __shfl_down_sync(FULL, v, d);
    if (lane + d < 32) v += t;
  }
  return v;
}

template <int L>
__global__ void __launch_bounds__(4 * L, (4 * L <= 320) ? 2 : 1) caar_fused_ldg_kernel(const KernelArgs A) {
  constexpr int NW = L / 8;  // warps per element
  __shared__ double tot[3][NW][16];

  const int t = threadIdx.x;
  const int lane = t & 31, w = t >> 5;
  const int r = t & 3;  // igp
  const size_t e = (size_t)(A.nets + blockIdx.x);
  constexpr size_t lf = (size_t)L * PTS;
  const size_t off = (size_t)t * 4;  // this thread's 4 points inside a scalar level-field

  // ---- issue the n0 loads first
  const Row dp = ld_row(A.dp3d + (e * A.ntl + A.n0) * lf + off);
  Row v1, v2;
  ld_row2(A.v + ((e * A.ntl + A.n0) * lf + off) * 2, v1, v2);
  const Row T = ld_row(A.T + (e * A.ntl + A.n0) * lf + off);
  Row Tv = T;
  if (A.qn0 != -1) Tv = ld_row(A.Qdp + ((e * A.qsize_d + 0) * 2 + A.qn0) * lf + off);  // holds Qdp for now

  // ---- per-thread constants
  double cx[4];  // cx[x] = Dvv[r^x][r]; static indices + selects keep Dvv in the constant bank
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    double c = A.dvv[(0 ^ x) * 4 + 0];
    if (r == 1) c = A.dvv[(1 ^ x) * 4 + 1];
    if (r == 2) c = A.dvv[(2 ^ x) * 4 + 2];
    if (r == 3) c = A.dvv[(3 ^ x) * 4 + 3];
    cx[x] = c;
  }
  Geo g;
  {
    const double* dinv = A.Dinv + e * 64 + r * 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Row q = ld_row(dinv + j * 4);
#pragma unroll
      for (int c = 0; c < 4; ++c) g.di[j][c] = q.x[c] * A.rrearth;
    }
  }
  const Row rmet = ld_row(A.rmetdet + e * 16 + r * 4);

  // ---- A: p = hyai0*ps0 + sum_{l<k} dp_l + dp_k/2   (PO:76-97)
  Row p;
#pragma unroll
  for (int j = 0; j < 4; ++j) p.x[j] = scan_down(dp.x[j], lane);
  if (lane >= 28) {
#pragma unroll
    for (int j = 0; j < 4; ++j) tot[0][w][r * 4 + j] = p.x[j];
  }
  __syncthreads();
  {
    double carry[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ww = 0; ww < NW - 1; ++ww)
      if (ww < w) {
#pragma unroll
        for (int j = 0; j < 4; ++j) carry[j] += tot[0][ww][r * 4 + j];
      }
    const double ptop = A.hyai0 * A.ps0;
#pragma unroll
    for (int j = 0; j < 4; ++j) p.x[j] = ptop + ((carry[j] + p.x[j]) - 0.5 * dp.x[j]);
  }
  Row rp;
#pragma unroll
  for (int j = 0; j < 4; ++j) rp.x[j] = 1.0 / p.x[j];

  // ---- B: grad_p, vgrad_p, vdp, vn0, divdp, vort (PO:101-124)
  Row gp0, gp1;
  gradient(p, g, cx, A.dvv, gp0, gp1);
  Row vgp;
#pragma unroll
  for (int j = 0; j < 4; ++j) vgp.x[j] = fma(v1.x[j], gp0.x[j], v2.x[j] * gp1.x[j]);

  Row divdp;
  {
    Row u, ww2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      u.x[j] = v1.x[j] * dp.x[j];
      ww2.x[j] = v2.x[j] * dp.x[j];
    }
    {  // derived_vn0 += eta_ave_w * vdp (PO:117-118)
      double* vn0 = A.vn0 + (e * lf + off) * 2;
      Row a0, a1;
      ld_row2(vn0, a0, a1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a0.x[j] = fma(A.eta_ave_w, u.x[j], a0.x[j]);
        a1.x[j] = fma(A.eta_ave_w, ww2.x[j], a1.x[j]);
      }
      st_row2(vn0, a0, a1);
    }
    // divergence_sphere (PO/sphere_operators.cpp:50-89); g.di carries the rrearth factor
    const Row met = ld_row(A.metdet + e * 16 + r * 4);
    Row gv0, gv1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gv0.x[j] = met.x[j] * fma(g.di[j][0], u.x[j], g.di[j][1] * ww2.x[j]);
      gv1.x[j] = met.x[j] * fma(g.di[j][2], u.x[j], g.di[j][3] * ww2.x[j]);
    }
    const Row dudx = deriv_i(gv0, cx);
    const Row dvdy = deriv_j(gv1, A.dvv);
#pragma unroll
    for (int j = 0; j < 4; ++j) divdp.x[j] = (dudx.x[j] + dvdy.x[j]) * rmet.x[j];
  }
  Row vort;
  {  // vorticity_sphere (PO/sphere_operators.cpp:91-129)
    const double* D = A.D + e * 64 + r * 16;
    Row vc0, vc1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const Row d = ld_row(D + j * 4);
      vc0.x[j] = fma(d.x[0], v1.x[j], d.x[2] * v2.x[j]);
      vc1.x[j] = fma(d.x[1], v1.x[j], d.x[3] * v2.x[j]);
    }
    const Row dvdx = deriv_i(vc1, cx);
    const Row dudy = deriv_j(vc0, A.dvv);
#pragma unroll
    for (int j = 0; j < 4; ++j) vort.x[j] = (dvdx.x[j] - dudy.x[j]) * rmet.x[j] * A.rrearth;
  }

  // ---- C: virtual temperature (PO:126-156)
  if (A.qn0 != -1) {
    const double c = A.Rwv / A.Rgas - 1.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) Tv.x[j] = T.x[j] * fma(c, Tv.x[j] / dp.x[j], 1.0);
  }

  // ---- D+E: both vertical integrals in scan form
  //   q_k = Rgas*T_v*dp/p ; phi_k = phis + sum_{l>k} q_l + q_k/2            (PO:280-312)
  //   omega_k = (vgrad_p - sum_{l<k} divdp_l - divdp_k/2) / p               (PO:314-352)
  Row q, sq, sd;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    q.x[j] = A.Rgas * Tv.x[j] * (dp.x[j] * rp.x[j]);
    sq.x[j] = scan_up(q.x[j], lane);
    sd.x[j] = scan_down(divdp.x[j], lane);
  }
  if (lane < 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) tot[1][w][r * 4 + j] = sq.x[j];
  }
  if (lane >= 28) {
#pragma unroll
    for (int j = 0; j < 4; ++j) tot[2][w][r * 4 + j] = sd.x[j];
  }
  __syncthreads();
  Row omega, phi;
  {
    double cq[4] = {0, 0, 0, 0}, cd[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) {
      if (ww > w) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cq[j] += tot[1][ww][r * 4 + j];
      }
      if (ww < w) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cd[j] += tot[2][ww][r * 4 + j];
      }
    }
    const Row phis = ld_row(A.phis + e * 16 + r * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      phi.x[j] = phis.x[j] + ((cq[j] + sq.x[j]) - 0.5 * q.x[j]);
      omega.x[j] = rp.x[j] * (vgp.x[j] - ((cd[j] + sd.x[j]) - 0.5 * divdp.x[j]));
    }
  }
  st_row(A.phi + e * lf + off, phi);
  {  // derived_omega_p += eta_ave_w * omega (PO:173). derived_eta_dot_dpdn += eta_ave_w*0 is value-neutral: skipped.
    double* op = A.omega_p + e * lf + off;
    Row a = ld_row(op);
#pragma unroll
    for (int j = 0; j < 4; ++j) a.x[j] = fma(A.eta_ave_w, omega.x[j], a.x[j]);
    st_row(op, a);
  }

  // ---- G: tendencies (PO:187-234)
  Row vt1, vt2, tt;
  {
    Row gT0, gT1;
    gradient(T, g, cx, A.dvv, gT0, gT1);
    const Row pec = ld_row(A.pecnd + e * lf + off);
    Row Ephi;
#pragma unroll
    for (int j = 0; j < 4; ++j) Ephi.x[j] = 0.5 * fma(v1.x[j], v1.x[j], v2.x[j] * v2.x[j]) + phi.x[j] + pec.x[j];
    Row gE0, gE1;
    gradient(Ephi, g, cx, A.dvv, gE0, gE1);
    const Row fcor = ld_row(A.fcor + e * 16 + r * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double vgT = fma(v1.x[j], gT0.x[j], v2.x[j] * gT1.x[j]);
      const double gl = A.Rgas * (Tv.x[j] * rp.x[j]);
      const double fv = fcor.x[j] + vort.x[j];
      vt1.x[j] = v2.x[j] * fv - gE0.x[j] - gl * gp0.x[j];
      vt2.x[j] = -v1.x[j] * fv - gE1.x[j] - gl * gp1.x[j];
      tt.x[j] = A.kappa * Tv.x[j] * omega.x[j] - vgT;
    }
  }

  // ---- H: apply (PO:236-257). Each thread reads nm1 and writes np1 only at its own points, after all of
  // its n0 reads: time levels may alias.
  {
    const Row mp = ld_row(A.spheremp + e * 16 + r * 4);
    const size_t onm1 = (e * A.ntl + A.nm1) * lf + off, onp1 = (e * A.ntl + A.np1) * lf + off;
    Row a0, a1;
    ld_row2(A.v + onm1 * 2, a0, a1);
    const Row Tm = ld_row(A.T + onm1);
    const Row dpm = ld_row(A.dp3d + onm1);
    Row oT, odp;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a0.x[j] = mp.x[j] * fma(A.dt2, vt1.x[j], a0.x[j]);
      a1.x[j] = mp.x[j] * fma(A.dt2, vt2.x[j], a1.x[j]);
      oT.x[j] = mp.x[j] * fma(A.dt2, tt.x[j], Tm.x[j]);
      odp.x[j] = mp.x[j] * fma(-A.dt2, divdp.x[j], dpm.x[j]);
    }
    st_row2(A.v + onp1 * 2, a0, a1);
    st_row(A.T + onp1, oT);
    st_row(A.dp3d + onp1, odp);
  }
}

template <int L>
cudaError_t launch_L(const KernelArgs& a, cudaStream_t s) {
  const int n = a.nete - a.nets;
  if (n <= 0) return cudaSuccess;
  caar_fused_ldg_kernel<L><<<n, 4 * L, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace

bool fused_ldg_supports(int nlev) { return nlev == 72 || nlev == 128 || nlev == 64 || nlev == 32 || nlev == 8 || nlev == 16; }

cudaError_t launch_fused_ldg(const KernelArgs& a, cudaStream_t s) {
  switch (a.nlev) {
    case 8: return launch_L<8>(a, s);
    case 16: return launch_L<16>(a, s);
    case 32: return launch_L<32>(a, s);
    case 64: return launch_L<64>(a, s);
    case 72: return launch_L<72>(a, s);
    case 128: return launch_L<128>(a, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace caar
