"""The N>1 host path on CPU with gloo, world_size 2: element partition, per-rank closed-form slices
(TestData.init_data(elem_offset)), per-rank run, all-reduce of the squared norms. The compute engine here is
the CPU oracle standing in for the GPU kernel (same element-local semantics); the GPU version of this flow is
bench.py under torchrun."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import harness
from tinman_sandbox_b200.partition import element_range, reduce_checksums, reduce_norms
from tinman_sandbox_b200.testdata import TestData

E_TOTAL = 9


def test_partition_covers_everything_once():
    for E in (1, 7, 10, 86400, 393216):
        for G in (1, 2, 3, 4, 8):
            if G > E:
                continue
            blocks = [element_range(g, G, E) for g in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == E
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        element_range(2, 2, 10)


CS_FIELDS = ("elem_state_dp3d", "elem_state_v", "elem_state_T", "elem_derived_eta_dot_dpdn", "elem_derived_omega_p",
             "elem_derived_phi", "elem_derived_vn0")


def _checksums(arrays, tl):
    """numpy stand-in for caar_checksums (include/caar_b200.h): sum, sum of squares, bit-pattern sum mod 2^64"""
    out = {"sum": [], "sumsq": [], "bits": [], "energy": [0.0, 0.0]}
    for n in CS_FIELDS:
        a = arrays[n][:, tl] if n.startswith("elem_state") else arrays[n]
        a = np.ascontiguousarray(a)
        out["sum"].append(a.sum())
        out["sumsq"].append((a * a).sum())
        out["bits"].append(a.view(np.uint64).sum(dtype=np.uint64))
    return out


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = element_range(rank, world, E_TOTAL)
    td = TestData(hi - lo).init_data(elem_offset=lo)
    orc = harness.PortOracle()
    s = harness.State(td.nelem, td.nlev)
    s.arrays, s.ctl, s.dt2, s.consts, s.dvv, s.ps0, s.hyai = td.arrays, td.ctl, td.dt2, td.consts, td.dvv, td.ps0, td.hyai
    orc.run(s, 2, 1)
    np1 = int(s.ctl[3])
    local = np.array([np.sum(s.arrays["elem_state_v"][:, np1] ** 2), np.sum(s.arrays["elem_state_T"][:, np1] ** 2),
                      np.sum(s.arrays["elem_state_dp3d"][:, np1] ** 2)])

    def allreduce(x):
        t = torch.from_numpy(x.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    norms = reduce_norms(local, allreduce)
    np.save(os.path.join(out_dir, f"norms_{rank}.npy"), norms)
    cs = reduce_checksums(_checksums(s.arrays, np1), allreduce)       # what caar_checksums returns per rank, from numpy
    np.save(os.path.join(out_dir, f"bits_{rank}.npy"), cs["bits"])
    np.save(os.path.join(out_dir, f"sums_{rank}.npy"), np.concatenate([cs["sum"], cs["sumsq"]]))
    np.save(os.path.join(out_dir, f"T_{rank}.npy"), s.arrays["elem_state_T"][:, np1])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_run(tmp_path):
    port = 29400 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    orc = harness.PortOracle()
    full = orc.init(E_TOTAL)
    orc.run(full, 2, 1)
    want = orc.norms(full)
    n0, n1 = np.load(tmp_path / "norms_0.npy"), np.load(tmp_path / "norms_1.npy")
    assert np.array_equal(n0, n1)                                   # every rank holds the reduced result
    assert np.max(np.abs(n0 - want) / want) < 1e-13                 # sum order differs from the Kahan loop
    # checksums: the 64-bit integer all-reduce of the bit-pattern sums is exact and partition-independent
    b0, b1 = np.load(tmp_path / "bits_0.npy"), np.load(tmp_path / "bits_1.npy")
    want_cs = _checksums(full.arrays, 1)
    assert np.array_equal(b0, b1) and np.array_equal(b0, np.array(want_cs["bits"], dtype=np.uint64))
    s0 = np.load(tmp_path / "sums_0.npy")
    want_s = np.array(want_cs["sum"] + want_cs["sumsq"])
    assert np.max(np.abs(s0 - want_s) / np.maximum(np.abs(want_s), 1e-300)) < 1e-13
    T = np.concatenate([np.load(tmp_path / "T_0.npy"), np.load(tmp_path / "T_1.npy")])
    assert np.array_equal(T, full.arrays["elem_state_T"][:, 1])     # slices == the global run, bit for bit
