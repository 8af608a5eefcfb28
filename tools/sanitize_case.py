#!/usr/bin/env python
"""A small pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
10-13 elements, nlev 72 / 128 / 24, both modes, both vertical coordinates, the host-array pipeline, the F90 relayout,
the tracer step, norms and saxpby. Prints 'sanitize_case done' at the end; results are not checked here (tests do)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinman_sandbox_b200 as tb  # noqa: E402
from tinman_sandbox_b200.testdata import TestData  # noqa: E402

for L, E in ((72, 10), (128, 7), (24, 5)):
    td = TestData(E, L, qsize_d=2).init_data()
    h = tb.Caar(E, L, 2)
    h.set_params(td.consts, td.dvv, td.ps0, td.hyai)
    h.set_control(*[int(x) for x in td.ctl], dt2=td.dt2)
    h.upload(td.arrays)
    for mode in (tb.MODE_FAST, tb.MODE_STRICT):
        h.set_vertical_coordinate(1)
        h.compute_and_apply_rhs(1, mode)
        h.set_vertical_coordinate(0, np.linspace(0.0, 1.0, L + 1))
        h.compute_and_apply_rhs(1, mode)
        h.set_vertical_coordinate(1)
        h.compute_and_apply_rhs_host(td.arrays, mode, 3)
        h.upload_vstar(np.ascontiguousarray(td.arrays["elem_derived_vn0"]))
        h.euler_step(0, 2, 10.0, mode)
    h.norms()
    h.download_f90({n: np.zeros_like(a) for n, a in td.arrays.items()}, names=None)
    h.close()
x, y = np.ones(5000), np.ones(5000)
tb.saxpby_host(3.0, 5.0, x, y, sweeps=2)
print("sanitize_case done")
