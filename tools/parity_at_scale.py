#!/usr/bin/env python
"""Per-field max relative error (max|a-b| / max|b|) of the fast kernels against the bit-exact reference-order kernel at
full size, after 1 and 10 calls. The reference-order kernel is itself checked bit-for-bit against the CPU reference on
element slices by tests/test_parity_gpu.py; this tool gives the numbers for the whole ne=120 / ne=256-per-GPU state.
    python tools/parity_at_scale.py [--nelem 86400] [--nlev 72] [--eulerian]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinman_sandbox_b200 as tb  # noqa: E402
from tinman_sandbox_b200.testdata import TestData  # noqa: E402


def run(td, ncalls, mode, eulerian):
    h = tb.Caar(td.nelem, td.nlev)
    h.set_params(td.consts, td.dvv, td.ps0, td.hyai)
    h.set_control(*[int(x) for x in td.ctl], dt2=td.dt2)
    if eulerian:
        h.set_vertical_coordinate(0, np.linspace(0.0, 1.0, td.nlev + 1))
    h.upload(td.arrays)
    h.compute_and_apply_rhs(ncalls, mode)
    out = {n: np.empty_like(td.arrays[n]) for n in tb.MUTATED_FIELDS}
    h.download(out, names=tb.MUTATED_FIELDS)
    h.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nelem", type=int, default=86400)
    ap.add_argument("--nlev", type=int, default=72)
    ap.add_argument("--eulerian", action="store_true")
    args = ap.parse_args()
    td = TestData(args.nelem, args.nlev).init_data()
    res = {"nelem": args.nelem, "nlev": args.nlev, "eulerian": args.eulerian, "calls": {}}
    for ncalls in (1, 10):
        a = run(td, ncalls, tb.MODE_STRICT, args.eulerian)
        b = run(td, ncalls, tb.MODE_FAST, args.eulerian)
        res["calls"][ncalls] = {n: float(np.max(np.abs(b[n] - a[n])) / max(np.max(np.abs(a[n])), 1e-300)) for n in a}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
