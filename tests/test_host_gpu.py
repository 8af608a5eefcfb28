"""The C++ host side on a GPU: the reference's own unmodified driver linked against the shim
(tinman_sandbox_b200/host/_dropin/pointers_only_b200) and the standalone driver (host/caar_driver) must print
the norms of the reference driver (tests/golden/pointers_only_stdout.txt)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tinman_sandbox_b200", "host")


def _norm_lines(txt):
    return [l for l in txt.splitlines() if "||" in l]


def _vals(txt):
    return np.array([float(x) for x in re.findall(r"=\s*([0-9.eE+-]+)", "\n".join(_norm_lines(txt)))])


def test_reference_driver_with_shim_strict_is_textually_identical(golden_dir, tmp_path):
    exe = os.path.join(HOST, "_dropin", "pointers_only_b200")
    if not os.path.exists(exe):
        pytest.skip("drop-in driver not built (needs the reference checkout at build time)")
    want = open(os.path.join(golden_dir, "pointers_only_stdout.txt")).read()
    out = subprocess.run([exe], capture_output=True, text=True, check=True, cwd=tmp_path,
                         env=dict(os.environ, CAAR_MODE="strict")).stdout
    got = "\n".join(l for l in out.splitlines() if "total time" not in l) + "\n"
    assert got == want
    # fast mode: same norms to 1e-13, and the dump files have the reference's format
    out = subprocess.run([exe, "--tinman-num-elems=4", "--tinman-num-exec=2", "--tinman-dump-res=yes"],
                         capture_output=True, text=True, check=True, cwd=tmp_path).stdout
    assert len(_vals(out)) == 6
    for f in ("elem_state_vx.txt", "elem_state_vy.txt", "elem_state_t.txt", "elem_state_dp3d.txt"):
        lines = open(tmp_path / f).read().splitlines()
        assert lines[0] == "[0, 0]" and len(lines) == 4 * 72 * 5


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_standalone_driver_prints_reference_norms(mode, golden_dir):
    exe = os.path.join(HOST, "caar_driver")
    want = _vals(open(os.path.join(golden_dir, "pointers_only_stdout.txt")).read())
    out = subprocess.run([exe, f"--caar-mode={mode}"], capture_output=True, text=True, check=True).stdout
    got = _vals(out)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want) / want) < 1e-13
    out2 = subprocess.run([exe, f"--caar-mode={mode}", "--caar-resident=no"], capture_output=True, text=True,
                          check=True).stdout
    assert np.max(np.abs(_vals(out2) - want) / want) < 1e-13


def test_standalone_driver_checksums_do_not_depend_on_the_partition():
    """--caar-checksums: per-GPU caar_checksums summed over ranks (NCCL when more than one GPU is visible). In strict
    mode the exact bit-pattern sums printed for 1 GPU equal the CPU reference's (bit-identical data), and a multi-GPU
    run — any partition — prints the same ones."""
    import torch
    from oracle import harness
    exe = os.path.join(HOST, "caar_driver")
    orc = harness.best_oracle(72)
    s = orc.init(10)
    orc.run(s)
    want = {}
    for name, short in (("elem_state_dp3d", "dp3d"), ("elem_state_v", "v"), ("elem_state_T", "T"),
                        ("elem_derived_eta_dot_dpdn", "eta_dot_dpdn"), ("elem_derived_omega_p", "omega_p"),
                        ("elem_derived_phi", "phi"), ("elem_derived_vn0", "vn0")):
        a = s.arrays[name][:, 1] if name.startswith("elem_state") and name != "elem_state_phis" else s.arrays[name]
        want[short] = format(int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64)), "016x")
    gpus = [1] + ([2] if torch.cuda.device_count() >= 2 else [])
    for g in gpus:
        out = subprocess.run([exe, "--caar-mode=strict", "--caar-checksums=yes", f"--caar-gpus={g}"], capture_output=True,
                             text=True, check=True).stdout
        got = {m.group(1): m.group(2) for m in re.finditer(r"^\s+(\w+)\s+\S+\s+\S+\s+([0-9a-f]{16})$", out, re.M)}
        assert got == want, (g, got, want)


def test_driver_cli_errors():
    exe = os.path.join(HOST, "caar_driver")
    assert subprocess.run([exe, "--tinman-num-elems=abc"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "--tinman-dump-res=maybe"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "--tinman-help"], capture_output=True).returncode == 0


@pytest.mark.parametrize("pulls", ["combined", "partial"])
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_hommexx_style_f90_pointer_interface(mode, pulls, golden_dir):
    """host/hommexx_shim.hpp (the twelve-argument Control::init, Derivative::init, Elements::init_2d /
    pull_from_f90_pointers / push_to_f90_pointers and the partial pull_* / push_* — the names and argument order of
    level_vectorized_ppscan/) driven by a C++ program that hands over Fortran-order arrays: the norms are the reference
    driver's."""
    exe = os.path.join(HOST, "hommexx_shim_test")
    want = _vals(open(os.path.join(golden_dir, "pointers_only_stdout.txt")).read())[3:]     # norms after the run
    # "partial": pull_3d / pull_4d / pull_eta_dot / pull_qdp and push_* instead of the combined calls
    out = subprocess.run([exe, "10", "1", mode, pulls], capture_output=True, text=True, check=True).stdout
    got = _vals(out)[:3]
    assert np.max(np.abs(got - want) / want) < 1e-13
    m = re.search(r"= ([0-9.eE+-]+) == Fortran T\(2,3,1,np1\) = ([0-9.eE+-]+)", out)
    assert m and m.group(1) == m.group(2)


def test_saxpby_driver_matches_reference_protocol():
    """host/saxpby_driver: the reference's saxpby_test driver protocol (I1 x 128 x 256 doubles, 100 sweeps of
    x = 3x + 5y, 'Init:' / 'saxpby:' timer lines) on HBM-resident arrays, values checked on the host."""
    exe = os.path.join(HOST, "saxpby_driver")
    p = subprocess.run([exe, "8", "--check"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert lines[0].split() == ["8", "128", "256"]
    assert lines[1].startswith("Init: ") and lines[2].startswith("saxpby: ") and "check: ok" in p.stdout
