"""Pins the parts of the CPU oracle that the plain C++ reference (pointers_only) does not cover — the F90 flat-pointer
boundary, preq_vertadv, the tracer-step operator, the weak-form / laplace operators — against the reference's OWN
HOMMEXX code: level_vectorized_ppscan ("lv") and tiled_vectorized_ppscan ("tv"), compiled unmodified from
/root/reference against the serial Kokkos stand-in oracle/kokkos_stub (oracle/Makefile -> oracle/_ref/libhommexx_*.so).

Findings these tests encode (reference defects, not ours):
  * vendored SIMD Vector: unary minus negates its argument IN PLACE (vector/KokkosKernels_Vector_SIMD.hpp:186-193), so
    one CaarFunctor call flips the sign of u(n0) in memory (CaarFunctor.hpp:139-142); the np1 results are right.
  * level_vectorized_ppscan/EulerStepFunctor.hpp does not compile (views typed [NUM_LEV][NP][NP] assigned from
    [NP][NP][NUM_LEV] subviews); tiled_vectorized_ppscan/EulerStepFunctor.hpp:58-59 writes v_buf(0|1, ilev, igp, jgp)
    into a view declared [NUM_LEV][2][NP][NP], i.e. with level and component swapped, and computes garbage.
    divergence_sphere_update itself (what the functor calls) is sound and is what pins caar_oracle_euler_step.
"""
import os
import subprocess

import numpy as np
import pytest

from oracle import harness as H

VARIANTS = ("lv", "tv")


def _hx(variant, nlev=72):
    if not H.HommexxOracle.available(variant, nlev):
        pytest.skip(f"oracle/_ref/libhommexx_{variant}_L{nlev}.so not built (needs /root/reference at build time)")
    return H.HommexxOracle(variant, nlev)


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("nlev", [72, 128])
def test_hommexx_caar_through_f90_pointers_matches_pointers_only(port, variant, nlev):
    """The whole path through the reference's own F90 boundary: Elements::init_2d / pull_from_f90_pointers,
    Control::init, Derivative::init, CaarFunctor per element, push_to_f90_pointers — on arrays converted with
    harness.to_f90 — agrees with pointers_only (the restatement is bit-identical to it, test_oracle.py) on every
    array the routine writes. HOMMEXX sums the scans in a different order (CaarFunctor.hpp:198-322): 1e-13."""
    hx = _hx(variant, nlev)
    for seed, random in ((0, False), (11, True)):
        s = port.init(3, nlev, qsize_d=hx.qsize_d)
        if random:
            H.randomize(s, seed=seed)
            s.arrays["elem_derived_eta_dot_dpdn"][...] = 0.0   # HOMMEXX overwrites it with 0 (CaarFunctor.hpp:169-179)
        want, got = s.copy(), s.copy()
        port.run(want)
        hx.run(got)
        n0, np1 = int(s.ctl[2]), int(s.ctl[3])
        for n in H.MUTATED:
            a, b = got.arrays[n], want.arrays[n]
            if n == "elem_state_v":                     # compare the level the routine writes; n0 is checked below
                a, b = a[:, np1], b[:, np1]
            assert rel(a, b) < 1e-13, (variant, n, rel(a, b))
        # the documented defect: u(n0) comes back negated, v(n0) and the third level untouched
        v_in, v_out = s.arrays["elem_state_v"], got.arrays["elem_state_v"]
        assert np.array_equal(v_out[:, n0, ..., 0], -v_in[:, n0, ..., 0])
        assert np.array_equal(v_out[:, n0, ..., 1], v_in[:, n0, ..., 1])
        assert np.array_equal(v_out[:, int(s.ctl[4])], v_in[:, int(s.ctl[4])])
        for n in ("elem_state_dp3d", "elem_state_T"):    # the other time levels survive the pull/push round trip
            assert np.array_equal(got.arrays[n][:, n0], s.arrays[n][:, n0])


@pytest.mark.parametrize("variant", VARIANTS)
def test_f90_pull_push_round_trip_is_exact(port, variant):
    """ncalls = 0: pull_from_f90_pointers followed by push_to_f90_pointers returns every array bit for bit, i.e.
    harness.to_f90 / from_f90 (and hence caar_upload_layout(CAAR_LAYOUT_F90), tested against them on the GPU) is the
    order the reference's own boundary code reads and writes."""
    hx = _hx(variant)
    s = H.randomize(port.init(2, 72, qsize_d=hx.qsize_d), seed=5)
    got = s.copy()
    hx.run(got, ncalls=0)
    for n in H.FIELD_NAMES:
        assert np.array_equal(got.arrays[n], s.arrays[n]), n


@pytest.mark.parametrize("variant", VARIANTS)
def test_strong_form_operators_against_hommexx(port, variant):
    hx = _hx(variant)
    rng = np.random.default_rng(3)
    L = 72
    s = H.randomize(port.init(2, L, qsize_d=hx.qsize_d), seed=7)
    sc, vec = rng.uniform(200, 300, (L, 4, 4)), rng.uniform(-40, 40, (L, 4, 4, 2))
    for ie in range(2):
        g = hx.sphere_op("gradient_sphere", sc, s, ie)
        assert np.array_equal(g, np.stack([port.gradient_sphere(sc[k], s, ie) for k in range(L)]))
        # HOMMEXX groups (1/metdet * rrearth) (SphereOperators.hpp:354-355): last-bit differences only
        d = hx.sphere_op("divergence_sphere", vec, s, ie)
        assert rel(d, np.stack([port.divergence_sphere(vec[k], s, ie) for k in range(L)])) < 1e-15
        w = hx.sphere_op("vorticity_sphere", vec, s, ie)
        assert rel(w, np.stack([port.vorticity_sphere(vec[k], s, ie) for k in range(L)])) < 1e-15


@pytest.mark.parametrize("variant", VARIANTS)
def test_weak_form_operators_bit_exact_against_hommexx(port, variant):
    """divergence_sphere_wk, laplace_simple, laplace_tensor, laplace_tensor_replace
    (level_vectorized_ppscan/SphereOperators.hpp:493-636, tiled_vectorized_ppscan/SphereOperators.hpp:421-549)."""
    hx = _hx(variant)
    rng = np.random.default_rng(4)
    E, L = 3, 72
    s = H.randomize(port.init(E, L, qsize_d=hx.qsize_d), seed=8)
    vin, sin = rng.uniform(-40, 40, (E, L, 4, 4, 2)), rng.uniform(200, 300, (E, L, 4, 4))
    tv = rng.uniform(-1, 1, (E, 4, 4, 2, 2))
    for name, f in (("divergence_sphere_wk", vin), ("laplace_simple", sin), ("laplace_tensor", sin)):
        po = port.sphere_wk(name, s, f, tv if name == "laplace_tensor" else None)
        for e in range(E):
            assert np.array_equal(hx.sphere_op(name, f[e], s, e, tensorvisc=tv[e]), po[e]), (name, e)
    po = port.sphere_wk("laplace_tensor", s, sin, tv)
    for e in range(E):
        assert np.array_equal(hx.sphere_op("laplace_tensor_replace", sin[e], s, e, tensorvisc=tv[e]), po[e])


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("nlev", [72, 128])
def test_preq_vertadv_bit_exact_against_hommexx(port, variant, nlev):
    """CaarFunctor::preq_vertadv (level_vectorized_ppscan/CaarFunctor.hpp:504-547) — the function the Eulerian branch
    of the restatement calls (caar_oracle.c rhs_element -> preq_vertadv)."""
    hx = _hx(variant, nlev)
    rng = np.random.default_rng(nlev)
    T, v = rng.uniform(200, 300, (nlev, 4, 4)), rng.uniform(-40, 40, (nlev, 4, 4, 2))
    eta, rp = rng.uniform(-1, 1, (nlev + 1, 4, 4)), 1.0 / rng.uniform(5, 15, (nlev, 4, 4))
    a, b = hx.preq_vertadv(T, v, eta, rp), port.preq_vertadv(T, v, eta, rp)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("variant", VARIANTS)
def test_tracer_step_operator_against_hommexx(port, variant):
    """caar_oracle_euler_step == the reference's divergence_sphere_update(alpha = -dt, beta = 1) applied to
    v_buf = vstar*qdp, q_buf = qdp (EulerStepFunctor.hpp:55-64): the operator AND the update are the reference's
    code; only the two products forming v_buf are restated. 1e-15: (1/metdet*rrearth) grouping."""
    hx = _hx(variant)
    rng = np.random.default_rng(6)
    E, L, Q = 2, 72, hx.qsize_d
    s = H.randomize(port.init(E, L, qsize_d=Q), seed=9)
    vstar = rng.uniform(-40, 40, (E, L, 4, 4, 2))
    dt, qn0 = 12.5, 1
    want = np.zeros((E, Q, L, 4, 4))
    port.euler_step(s, vstar, want, qn0, Q, dt)
    for e in range(E):
        for iq in range(Q):
            q = s.arrays["elem_state_Qdp"][e, iq, qn0]
            got = hx.sphere_op("divergence_sphere_update", np.ascontiguousarray(vstar[e] * q[..., None]), s, e,
                               alpha=-dt, beta=1.0, out=q)
            assert rel(got, want[e, iq]) < 1e-15


def test_tv_euler_step_functor_defect_is_what_we_say_it_is(port):
    """tiled_vectorized_ppscan/EulerStepFunctor.hpp:58-59 indexes v_buf (declared [NUM_LEV][2][NP][NP]) as
    (comp, lev, igp, jgp). Emulating exactly that store/load pattern reproduces the functor's output bit for bit —
    so the functor as written is not an oracle, and the mismatch is its, not the restatement's."""
    hx = _hx("tv")
    rng = np.random.default_rng(3)
    E, L = 2, 72
    s = H.randomize(port.init(E, L, qsize_d=hx.qsize_d), seed=9)
    vstar = rng.uniform(-40, 40, (E, L, 4, 4, 2))
    dt, qn0, ie, iq = 12.5, 1, 1, 1
    functor = hx.euler_step(s, vstar, qn0, 2, dt)
    q = s.arrays["elem_state_Qdp"][ie, iq, qn0]
    qh = q.transpose(0, 2, 1).reshape(L, 16)
    vh = vstar[ie].transpose(0, 3, 2, 1).reshape(L, 2, 16)
    buf, as_read = np.zeros(L * 32), np.zeros((L, 2, 16))
    for lev in range(L):                                  # the serial TeamThreadRange order
        buf[lev * 16:(lev + 1) * 16] = vh[lev, 0] * qh[lev]              # v_buf(0, lev, ., .)
        buf[32 + lev * 16:32 + (lev + 1) * 16] = vh[lev, 1] * qh[lev]    # v_buf(1, lev, ., .)
        as_read[lev, 0] = buf[lev * 32:lev * 32 + 16]                    # v(lev, 0, ., .)
        as_read[lev, 1] = buf[lev * 32 + 16:lev * 32 + 32]               # v(lev, 1, ., .)
    v_po = np.ascontiguousarray(as_read.reshape(L, 2, 4, 4).transpose(0, 3, 2, 1))
    emulated = hx.sphere_op("divergence_sphere_update", v_po, s, ie, alpha=-dt, beta=1.0, out=q)
    assert np.array_equal(emulated, functor[ie, iq])
    want = np.zeros_like(functor)
    port.euler_step(s, vstar, want, qn0, 2, dt)
    assert rel(functor[ie, iq], want[ie, iq]) > 1e-4      # the defect is visible


@pytest.mark.skipif(not os.path.isdir("/root/reference/compute_and_apply_rhs_test"), reason="needs /root/reference")
def test_lv_euler_step_functor_does_not_compile():
    """level_vectorized_ppscan/EulerStepFunctor.hpp:47-49 assigns [NP][NP][NUM_LEV] subviews to views typed
    [NUM_LEV][NP][NP]: rejected by the compile-time extent check (Kokkos' ViewMapping::is_assignable; same rule in the
    stand-in). Nothing in the reference's build includes this header (level_vectorized_ppscan/CMakeLists.txt:21-28)."""
    root = os.path.join(os.path.dirname(os.path.abspath(H.__file__)))
    lv = "/root/reference/compute_and_apply_rhs_test/cxx/level_vectorized_ppscan"
    cfg = os.path.join(root, "_ref", "hxcfg72")
    if not os.path.exists(os.path.join(cfg, "config.h.c")):
        pytest.skip("oracle/_ref/hxcfg72 not generated")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-I" + os.path.join(root, "kokkos_stub"), "-I" + cfg,
                        "-I" + lv, "-x", "c++", os.path.join(lv, "EulerStepFunctor.hpp")], capture_output=True, text=True)
    assert r.returncode != 0 and "Incompatible View copy construction" in r.stderr


@pytest.mark.parametrize("nlev", [72, 128])
@pytest.mark.parametrize("moist", [True, False])
def test_eulerian_restatement_against_reference_parts_plus_the_fortran_lines(port, nlev, moist):
    """Double entry for the one piece of the path nothing here can execute (rsplit == 0 of
    fortran/routine_extracted.F90): the Eulerian output is rebuilt from code the REFERENCE executes — the pointers_only
    build for everything the two branches share (Lagrangian results), its divergence_sphere for div(v dp), HOMMEXX's
    preq_vertadv (level_vectorized_ppscan/CaarFunctor.hpp:504-547) — plus an independent numpy transliteration of the
    remaining Fortran lines (F:233-254 eta_dot_dpdn from hybi, F:268-275 accumulation, F:325-334 the -T_vadv / -v_vadv
    terms, F:515-517 the dp3d update), and compared with the C restatement's Eulerian branch. T, v, dp3d(np1) are
    affine in the tendencies, so  X_eul(np1) = X_lag(np1) - spheremp*dt2*X_vadv  up to rounding."""
    if not H.ref_available(nlev):
        pytest.skip("oracle/_ref not built")
    hx, ref = _hx("lv", nlev), H.RefOracle(nlev)
    E, L = 3, nlev
    base = H.randomize(port.init(E, L), seed=100 + nlev)
    if not moist:
        base.ctl[5] = -1
    hybi = np.sort(np.random.default_rng(7).uniform(0.0, 1.0, L + 1))
    hybi[0], hybi[L] = 0.0, 1.0
    lag, eul = base.copy(), base.copy()
    ref.run(lag)                     # the reference's own build: the vertically Lagrangian branch
    port.run_eulerian(eul, hybi)     # the restatement under test
    A = base.arrays
    n0, np1, dt2, eta_ave_w = int(base.ctl[2]), int(base.ctl[3]), float(base.dt2), float(base.consts[1])
    for ie in range(E):
        dp, v, T = A["elem_state_dp3d"][ie, n0], A["elem_state_v"][ie, n0], A["elem_state_T"][ie, n0]
        mp = A["elem_spheremp"][ie]
        # divdp(:,:,k) = divergence_sphere(v*dp)   (F:176-186; the reference's operator)
        divdp = np.stack([ref.divergence_sphere(v[k] * dp[k][..., None], base, ie) for k in range(L)])
        # F:233-254
        eta = np.zeros((L + 1, 4, 4))
        sdot_sum = np.zeros((4, 4))
        for k in range(L):
            sdot_sum = sdot_sum + divdp[k]
            eta[k + 1] = sdot_sum
        for k in range(L - 1):
            eta[k + 1] = hybi[k + 1] * sdot_sum - eta[k + 1]
        eta[0] = 0.0
        eta[L] = 0.0
        # F:259-260, executed by the reference's C++
        T_vadv, v_vadv = hx.preq_vertadv(T, v, eta, 1.0 / dp)
        want = {
            "elem_state_T": lag.arrays["elem_state_T"][ie, np1] - mp * dt2 * T_vadv,                       # F:333, 514
            "elem_state_v": lag.arrays["elem_state_v"][ie, np1] - (mp * dt2)[..., None] * v_vadv,           # F:325-331, 512-513
            "elem_state_dp3d": lag.arrays["elem_state_dp3d"][ie, np1] - mp * dt2 * (eta[1:] - eta[:-1]),    # F:515-517
        }
        for n, w in want.items():
            assert rel(eul.arrays[n][ie, np1], w) < 1e-13, (n, ie)
        w_eta = A["elem_derived_eta_dot_dpdn"][ie] + eta_ave_w * eta                                      # F:268-275
        assert rel(eul.arrays["elem_derived_eta_dot_dpdn"][ie], w_eta) < 1e-14
        for n in ("elem_derived_phi", "elem_derived_omega_p", "elem_derived_vn0"):                         # no vertical-flux term
            assert np.array_equal(eul.arrays[n][ie], lag.arrays[n][ie]), n
        # the vertical terms are not a rounding-level correction: the comparison above has teeth
        assert rel(eul.arrays["elem_state_T"][ie, np1], lag.arrays["elem_state_T"][ie, np1]) > 1e-6
