"""Element partition across ranks — the reference's own nets/nete hook (pointers_only/data_structures.hpp:58-66,
compute_and_apply_rhs.cpp:65-74) applied per GPU: rank g of G owns the contiguous block
[floor(g*E/G), floor((g+1)*E/G)). compute_and_apply_rhs has no inter-element coupling, so no data-path
collective exists; the only exchange of a job is the final sum of the three squared norms."""
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
