#!/usr/bin/env python
"""BASELINE configs[1]: saxpby FP64 sweep 1 MB - 16 GB on one B200 as the HBM-bandwidth calibration.

x = a*x + b*y with a=3, b=5 (saxpby_test/cxx/main.cpp:39-41) over n doubles per array, total footprint
2*n*8 bytes, 24 bytes of traffic per element per sweep (read x, read y, write x). Device memory comes from
torch, the kernel is the library's own (caar_saxpby_device through the C-ABI), timing = CUDA events on the
launching stream. Writes one JSON document (default profiles/r1_saxpby_sweep.json).
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinman_sandbox_b200 as tb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_saxpby_sweep.json"))
    ap.add_argument("--max-gb", type=float, default=16.0)
    ap.add_argument("--min-gb", type=float, default=0.0, help="first footprint of the sweep (0 = 1 MB)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the reference's OpenMP saxpby on the host")
    args = ap.parse_args()
    lib = tb.load_library()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    rows = []
    total = 1 << 20  # bytes over both arrays
    if args.min_gb > 0:
        total = int(args.min_gb * (1 << 30))
    while total <= args.max_gb * (1 << 30):
        n = total // 16
        x = torch.ones(n, dtype=torch.float64, device=dev)
        y = torch.full((n,), 1e-3, dtype=torch.float64, device=dev)
        sweeps = max(5, min(200, int(4e9 // (24 * n))))

        def run(k):
            for _ in range(k):
                rc = lib.caar_saxpby_device(0.5, 5.0, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n,
                                            C.c_void_p(stream.cuda_stream))
                assert rc == 0, lib.caar_last_error()
        run(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(sweeps)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / sweeps
        rows.append({"footprint_bytes": total, "n": n, "sweeps": sweeps, "ms_per_sweep": ms,
                     "GBps": 24.0 * n / (ms * 1e-3) / 1e9})
        print(f"{total / 2**20:10.0f} MiB  {rows[-1]['GBps']:8.1f} GB/s", flush=True)
        del x, y
        total *= 2
    big = [r["GBps"] for r in rows if r["footprint_bytes"] >= (1 << 30)]
    doc = {"kernel": "saxpby_kernel (x=a*x+b*y, FP64)", "bytes_per_element": 24,
           "hbm_plateau_GBps": max(big) if big else None, "gpu": torch.cuda.get_device_name(0), "rows": rows}
    # CPU leg: the reference's own saxpby (saxpby_test/cxx/common.cpp, OpenMP) on the host cores,
    # at the reference's default size I1=1000 (262 MB per array), 100 sweeps like saxpby_test/cxx/main.cpp:39-41
    try:
        if args.no_cpu:
            raise RuntimeError("skipped (--no-cpu)")
        import numpy as np
        from oracle import harness
        ref = harness.RefSaxpby()
        i1 = 1000
        n = i1 * 128 * 256
        xh, yh = np.ones(n), np.full(n, 1e-3)
        ref.run(0.5, 5.0, xh, yh, sweeps=2)
        sec = ref.run(0.5, 5.0, xh, yh, sweeps=100)
        doc["cpu_reference"] = {"GBps": 24.0 * n * 100 / sec / 1e9, "I1": i1, "sweeps": 100,
                                "threads": len(os.sched_getaffinity(0)), "kind": "reference (OpenMP)"}
        print(f"CPU reference saxpby: {doc['cpu_reference']['GBps']:.1f} GB/s on {doc['cpu_reference']['threads']} threads")
    except Exception as exc:  # the CPU leg is optional (oracle/_ref may be absent)
        doc["cpu_reference"] = {"unavailable": str(exc)}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
