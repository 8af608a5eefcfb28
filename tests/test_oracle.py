"""Pins the CPU oracle (oracle/caar_oracle.c) before anything is compared with it:

  1. the reference's own golden vectors (fortran/test_mod.F90, committed as tests/golden/fortran_golden.npz),
  2. the norms the unmodified reference driver prints (tests/golden/pointers_only_stdout.txt),
  3. the real reference compiled from /root/reference (oracle/_ref/*.so), bit for bit, when present,
  4. committed outputs of the real reference (tests/golden/ref_outputs_E3.npz) — always.
"""
import os
import re

import numpy as np
import pytest

from oracle import harness


def test_golden_fortran_vectors(port, golden_dir):
    """Reference KAT: fortran/main.F90:241-274 compares T, v1, v2 of element 1 at np1 with
    Ttest/v1test/v2test. The Fortran driver fills Dvv from single-precision literals
    (fortran/main.F90:79-90), so Dvv is rounded through float32 here."""
    g = np.load(os.path.join(golden_dir, "fortran_golden.npz"))
    s = port.init(3)                                # nelemd = 3 is the Fortran default (kinds.F90:21)
    s.dvv[...] = s.dvv.astype(np.float32).astype(np.float64)
    port.run(s)
    np1 = int(s.ctl[3])
    # golden order: i fastest, then j, then k  <->  C [k][igp=i][jgp=j]
    T = s.arrays["elem_state_T"][0, np1].transpose(0, 2, 1).reshape(-1)
    v1 = s.arrays["elem_state_v"][0, np1, ..., 0].transpose(0, 2, 1).reshape(-1)
    v2 = s.arrays["elem_state_v"][0, np1, ..., 1].transpose(0, 2, 1).reshape(-1)
    assert np.array_equal(T, g["Ttest"])            # 17 significant digits printed: bit-exact
    # v literals carry 15 significant digits
    assert np.max(np.abs(v1 - g["v1test"]) / np.abs(g["v1test"])) < 1e-14
    assert np.max(np.abs(v2 - g["v2test"]) / np.abs(g["v2test"])) < 1e-14


def test_golden_fortran_vectors_double_dvv(port, golden_dir):
    """With the C++ drivers' double-precision Dvv the same vectors are matched to the float rounding of Dvv."""
    g = np.load(os.path.join(golden_dir, "fortran_golden.npz"))
    s = port.init(3)
    port.run(s)
    T = s.arrays["elem_state_T"][0, 1].transpose(0, 2, 1).reshape(-1)
    assert np.max(np.abs(T - g["Ttest"]) / np.abs(g["Ttest"])) < 1e-11


def test_printed_norms_of_reference_driver(port, golden_dir):
    """pointers_only/main.cpp:105,131 print ||v||,||T||,||dp|| before and after one call on 10 elements."""
    txt = open(os.path.join(golden_dir, "pointers_only_stdout.txt")).read()
    vals = [float(x) for x in re.findall(r"=\s*([0-9.eE+-]+)", txt)]
    assert len(vals) == 6
    s = port.init(10)
    n0 = port.norms(s)
    port.run(s)
    n1 = port.norms(s)
    assert np.array_equal(n0, np.array(vals[:3]))
    assert np.array_equal(n1, np.array(vals[3:]))


def test_committed_reference_outputs(port, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_outputs_E3.npz"))
    s = port.init(3)
    for call in (1, 2):
        port.run(s)
        for n in harness.MUTATED:
            a = s.arrays[n]
            if n in ("elem_state_dp3d", "elem_state_v", "elem_state_T"):
                a = a[:, int(s.ctl[3])]
            assert np.array_equal(a, g[f"call{call}_{n}"]), (call, n)


@pytest.mark.parametrize("nlev", [72, 128])
def test_port_bit_identical_to_compiled_reference(port, nlev):
    if not harness.ref_available(nlev):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ro = harness.RefOracle(nlev)
    a, b = port.init(4, nlev), ro.init(4)
    for n in harness.FIELD_NAMES:
        assert np.array_equal(a.arrays[n], b.arrays[n]), n
    assert np.array_equal(a.dvv, b.dvv) and np.array_equal(a.hyai, b.hyai)
    assert np.array_equal(a.consts, b.consts) and np.array_equal(a.ctl, b.ctl)
    port.run(a, 2, 1)
    ro.run(b, 2, 1)
    for n in harness.FIELD_NAMES:
        assert np.array_equal(a.arrays[n], b.arrays[n]), n
    assert np.array_equal(port.norms(a), ro.norms(b))
    # random geometry/fields, dry and moist, a rotated set of time levels, threads on both sides
    for qn0, tls in ((0, (0, 1, 2)), (-1, (2, 0, 1)), (1, (1, 2, 0))):
        a = harness.randomize(port.init(5, nlev), seed=7 + qn0)
        a.ctl[2:5] = tls
        a.ctl[5] = qn0
        b = a.copy()
        port.run(a, 1, 2)
        ro.run(b, 1, 3)
        for n in harness.FIELD_NAMES:
            assert np.array_equal(a.arrays[n], b.arrays[n]), (qn0, n)


def test_inputs_untouched_and_partition(port):
    """Only the arrays of PO/compute_and_apply_rhs.cpp:117-118,172-173,251-254 (+phi) change, only at np1,
    only inside [nets,nete)."""
    s = harness.randomize(port.init(6))
    ref = s.copy()
    s.ctl[0], s.ctl[1] = 2, 5
    port.run(s)
    for n in harness.FIELD_NAMES:
        if n not in harness.MUTATED:
            assert np.array_equal(s.arrays[n], ref.arrays[n]), n
    for n in harness.MUTATED:
        a, b = s.arrays[n], ref.arrays[n]
        assert np.array_equal(a[:2], b[:2]) and np.array_equal(a[5:], b[5:]), n
    for n in ("elem_state_dp3d", "elem_state_v", "elem_state_T"):
        assert np.array_equal(s.arrays[n][:, 0], ref.arrays[n][:, 0])
        assert np.array_equal(s.arrays[n][:, 2], ref.arrays[n][:, 2])
        assert not np.array_equal(s.arrays[n][2:5, 1], ref.arrays[n][2:5, 1])
    # running [0,6) in one go == running three disjoint ranges
    full = ref.copy()
    port.run(full)
    parts = ref.copy()
    for lo, hi in ((0, 1), (1, 4), (4, 6)):
        parts.ctl[0], parts.ctl[1] = lo, hi
        port.run(parts)
    for n in harness.MUTATED:
        assert np.array_equal(full.arrays[n], parts.arrays[n]), n


def test_saxpby_port():
    po = harness.PortOracle()
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(1 << 16), rng.standard_normal(1 << 16)
    want = x.copy()
    for _ in range(3):
        want = 3.0 * want + 5.0 * y
    po.saxpby(3.0, 5.0, x, y, sweeps=3, nthreads=2)
    assert np.array_equal(x, want)
    if os.path.exists(os.path.join(harness.REF_DIR, "libsaxpby_ref.so")):
        x2 = rng.standard_normal(2 * 128 * 256)
        y2 = rng.standard_normal(2 * 128 * 256)
        w2 = 3.0 * x2 + 5.0 * y2
        harness.RefSaxpby().run(3.0, 5.0, x2, y2, 1)
        assert np.array_equal(x2, w2)


def test_f90_layout_convention_against_fortran_golden(port, golden_dir):
    """The Fortran memory order used for the F90 boundary (harness.to_f90) is the one the reference's Fortran
    KAT is written in: T(:,:,:,np1) and v(:,:,1:2,:,np1) of element 1, flattened, ARE Ttest / v1test / v2test
    (fortran/test_mod.F90), with no transposition left to do."""
    g = np.load(os.path.join(golden_dir, "fortran_golden.npz"))
    s = port.init(3)
    s.dvv[...] = s.dvv.astype(np.float32).astype(np.float64)
    port.run(s)
    f = harness.to_f90({n: s.arrays[n] for n in ("elem_state_T", "elem_state_v")})
    assert np.array_equal(f["elem_state_T"][0, 1].reshape(-1), g["Ttest"])
    v = f["elem_state_v"][0, 1]                      # [lev][c][j][i]
    for c, key in ((0, "v1test"), (1, "v2test")):
        got = v[:, c].reshape(-1)
        assert np.max(np.abs(got - g[key])) / np.max(np.abs(g[key])) < 1e-14
    back = harness.from_f90(harness.to_f90(s.arrays))
    for n in harness.FIELD_NAMES:
        assert np.array_equal(back[n], s.arrays[n]), n


def test_divergence_operator_pinned_to_reference(port):
    """The operator the tracer-step oracle is built on: our restatement of divergence_sphere equals the
    reference's own function (oracle/_ref) bit for bit on random geometry and fields."""
    if not harness.ref_available(72):
        pytest.skip("oracle/_ref not built")
    ref = harness.RefOracle(72)
    s = harness.randomize(port.init(5), seed=44)
    rng = np.random.default_rng(8)
    for ie in range(5):
        v = rng.uniform(-30, 30, size=(4, 4, 2))
        assert np.array_equal(port.divergence_sphere(v, s, ie), ref.divergence_sphere(v, s, ie))


def test_euler_step_oracle_properties(port):
    """Tracer step oracle (LV/EulerStepFunctor.hpp:33-66): dt = 0 returns Qdp(qn0); linear in Qdp; a constant
    mixing ratio with non-divergent flow... is not testable without DSS, so linearity + the operator pin above."""
    s = harness.randomize(port.init(4, 24, 3), seed=5)
    rng = np.random.default_rng(1)
    vstar = rng.uniform(-20, 20, size=(4, 24, 4, 4, 2))
    q0 = np.zeros((4, 3, 24, 4, 4))
    port.euler_step(s, vstar, q0, qn0=1, qsize=3, dt=0.0)
    assert np.array_equal(q0, s.arrays["elem_state_Qdp"][:, :, 1])
    qa = np.zeros_like(q0)
    port.euler_step(s, vstar, qa, qn0=0, qsize=2, dt=300.0)
    assert np.all(qa[:, 2] == 0.0)                       # tracers >= qsize untouched
    s2 = s.copy()
    s2.arrays["elem_state_Qdp"] *= 2.0
    qb = np.zeros_like(q0)
    port.euler_step(s2, vstar, qb, qn0=0, qsize=2, dt=300.0)
    assert np.array_equal(qb, 2.0 * qa)                  # exact: scaling by 2 commutes with every rounding


def test_eulerian_branch_oracle_properties(port):
    """rsplit == 0 restatement (fortran/routine_extracted.F90:227-262; PARITY UNPINNED): the properties the
    Fortran formulas imply. (1) the vertical flux telescopes (eta = 0 at top and bottom), so the column sum of
    dp3d(np1)/spheremp equals the Lagrangian branch's; (2) with vertically uniform T and v the vertical
    advection vanishes and T, v come out as on the Lagrangian branch; (3) hybi only enters through eta."""
    L = 24
    base = harness.randomize(port.init(3, L), seed=17)
    hybi = np.linspace(0.0, 1.0, L + 1) ** 2
    lag, eul = base.copy(), base.copy()
    port.run(lag)
    port.run_eulerian(eul, hybi)
    mp = base.arrays["elem_spheremp"][:, None]
    col_l = (lag.arrays["elem_state_dp3d"][:, 1] / mp).sum(axis=1)
    col_e = (eul.arrays["elem_state_dp3d"][:, 1] / mp).sum(axis=1)
    assert np.max(np.abs(col_l - col_e) / np.abs(col_l)) < 1e-12
    assert not np.array_equal(lag.arrays["elem_state_dp3d"], eul.arrays["elem_state_dp3d"])
    assert not np.array_equal(lag.arrays["elem_state_T"], eul.arrays["elem_state_T"])
    for n in ("elem_derived_phi", "elem_derived_omega_p", "elem_derived_vn0"):       # no vertical-flux term in these
        assert np.array_equal(lag.arrays[n], eul.arrays[n]), n
    # interfaces 0 and L carry no flux
    d_eta = eul.arrays["elem_derived_eta_dot_dpdn"] - base.arrays["elem_derived_eta_dot_dpdn"]
    assert np.all(d_eta[:, 0] == 0) and np.all(d_eta[:, L] == 0) and np.any(d_eta[:, 1:L] != 0)
    # vertically uniform T and v: no vertical advection
    uni = base.copy()
    uni.arrays["elem_state_T"][:, :, :] = uni.arrays["elem_state_T"][:, :, :1]
    uni.arrays["elem_state_v"][:, :, :] = uni.arrays["elem_state_v"][:, :, :1]
    lag2, eul2 = uni.copy(), uni.copy()
    port.run(lag2)
    port.run_eulerian(eul2, hybi)
    assert np.array_equal(lag2.arrays["elem_state_T"], eul2.arrays["elem_state_T"])
    assert np.array_equal(lag2.arrays["elem_state_v"], eul2.arrays["elem_state_v"])
