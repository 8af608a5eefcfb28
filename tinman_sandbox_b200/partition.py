"""Element partition across ranks — the reference's own nets/nete hook (pointers_only/data_structures.hpp:58-66,
compute_and_apply_rhs.cpp:65-74) applied per GPU: rank g of G owns the contiguous block
[floor(g*E/G), floor((g+1)*E/G)). compute_and_apply_rhs has no inter-element coupling, so no data-path
collective exists; the only exchanges of a job are the final sums of the three squared norms and of the checksums."""
from __future__ import annotations

import numpy as np


def element_range(rank: int, world: int, nelem: int):
    if not (0 <= rank < world) or nelem < 0:
        raise ValueError("bad partition arguments")
    return (nelem * rank) // world, (nelem * (rank + 1)) // world


def reduce_norms(local_sumsq, all_reduce_sum=None):
    """local sums of squares (v, T, dp3d) -> global 2-norms. all_reduce_sum(np.ndarray) -> np.ndarray sums over
    ranks (torch.distributed with NCCL on GPUs, gloo in the CPU tests); None = single rank."""
    s = np.asarray(local_sumsq, dtype=np.float64)
    if all_reduce_sum is not None:
        s = all_reduce_sum(s)
    return np.sqrt(s)


def reduce_checksums(local, all_reduce_sum=None):
    """Per-rank caar_checksums (dict with float64 `sum`, `sumsq`, `energy` and uint64 `bits`) -> the job's: the doubles are
    summed as doubles, the bit-pattern sums as 64-bit integers with wrap-around (exact, independent of the partition).
    all_reduce_sum(np.ndarray) -> np.ndarray sums over ranks for float64 AND int64 arrays; None = single rank."""
    f = np.concatenate([np.asarray(local["sum"], dtype=np.float64), np.asarray(local["sumsq"], dtype=np.float64),
                        np.asarray(local["energy"], dtype=np.float64)])
    b = np.asarray(local["bits"], dtype=np.uint64).view(np.int64).copy()
    if all_reduce_sum is not None:
        f = all_reduce_sum(f)
        b = all_reduce_sum(b)
    n = len(local["sum"])
    return {"sum": f[:n], "sumsq": f[n:2 * n], "energy": f[2 * n:], "bits": b.view(np.uint64)}
