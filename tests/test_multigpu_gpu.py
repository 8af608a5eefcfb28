"""The N > 1 path on real GPUs (skipped on a one-GPU box): bench.py under torchrun with 2 ranks over NCCL.

What makes a multi-GPU run self-validating ("nothing verifies the multi-GPU result"):
  * every rank checks a sample of its own slice against the CPU reference and the line carries parity.ok;
  * the bit-pattern checksums of all seven mutated arrays, all-reduced as 64-bit integers, are EXACT and do not depend
    on the partition: in strict mode a 2-rank strong-scaled run must print the very same 7 numbers as the 1-rank run
    of the same elements — and those equal the CPU reference's (tests/test_parity_gpu.py::test_checksums_against_the_oracle).
"""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _bench(gpus, extra):
    args = ["--gpus", str(gpus), "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "1", "--no-clock-topup"] + extra
    if gpus == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py")] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py")] + args
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_two_ranks_agree_with_one_rank_bit_for_bit():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    E = 4001                                      # odd: the two slices differ in size
    one = _bench(1, ["--nelem", str(E), "--mode", "strict"])
    two = _bench(2, ["--nelem", str(E), "--mode", "strict", "--scaling", "strong"])
    assert one["parity"]["ok"] and two["parity"]["ok"]
    assert one["parity"]["max_rel_err"] == 0.0 and two["parity"]["max_rel_err"] == 0.0      # strict: bit-exact
    assert two["n_gpus"] == 2 and two["config"]["elements_per_gpu"] in (2000, 2001)
    assert one["checksums"]["calls"] == two["checksums"]["calls"]
    assert one["checksums"]["bits"] == two["checksums"]["bits"]
    for a, b in zip(one["norms_np1"], two["norms_np1"]):
        assert abs(a - b) <= 1e-13 * abs(a)
    for k in ("kinetic", "internal"):
        assert abs(one["checksums"]["energy"][k] - two["checksums"]["energy"][k]) <= 1e-13 * abs(one["checksums"]["energy"][k])


def test_two_ranks_weak_scaling_fast_mode_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = _bench(2, ["--nelem", "5400"])
    assert d["parity"]["ok"] and d["parity"]["elements_checked"] == 128 and d["parity"]["max_rel_err"] <= 1e-12
    assert d["parity"]["e2e_max_rel_err"] <= 1e-12 and d["parity"]["norms_rel_err"] <= 1e-13
    assert d["scaling"] == "weak" and d["gpu_launches"] > 0 and d["e2e"]["link"]["frac"] > 0


def test_single_rank_bench_line_is_self_validating():
    d = _bench(1, ["--nelem", "2000"])
    assert d["parity"]["ok"] and d["parity"]["oracle"] in ("reference", "port")
    assert len(d["checksums"]["bits"]) == 7 and d["resident_protocol"]["value"] > 0
    assert d["clocks"] is None or d["clocks"].get("samples", 0) >= 5 or "rejected" in d["clocks"]
    assert d["roofline"]["frac"] > 0 and d["e2e"]["link"]["duplex_peak_gbs"] > 0
