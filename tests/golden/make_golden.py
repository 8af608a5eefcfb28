#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerates the committed golden fixtures. Runs only where
/root/reference exists (the build container); the fixtures it writes travel with the repo.

  fortran_golden.npz        Ttest / v1test / v2test, the reference's own known-answer vectors
                            (compute_and_apply_rhs_test/fortran/test_mod.F90:8-296, 299-591, 594-886):
                            T, v(:,:,1,:), v(:,:,2,:) at time level np1 of element 1 after
                            compute_and_apply_rhs, 1152 = np*np*nlev values each, stored in Fortran
                            order (i fastest, then j, then k) exactly as main.F90:241-252 reads them.
                            The Fortran driver initialises Dvv from SINGLE-precision literals
                            (main.F90:79-90), so the vectors correspond to Dvv = (double)(float)Dvv.
  pointers_only_stdout.txt  stdout of the unmodified reference driver (pointers_only/main.cpp) built by
                            oracle/Makefile, default config (10 elements, 1 call): the printed-norm KAT.
  ref_outputs_E3.npz        every array the reference routine mutates, for 3 elements of the closed-form
                            init after 1 and after 2 calls (oracle/_ref/libcaar_ref_L72.so), so the GPU
                            tests can check against the REAL reference even if oracle/_ref is absent.
"""
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/compute_and_apply_rhs_test"


def parse_fortran_vectors(path):
    txt = open(path).read()
    out = {}
    for name in ("Ttest", "v1test", "v2test"):
        m = re.search(name + r"\(np\*np\*nlev\)\s*=\s*\(/(.*?)/\)", txt, re.S)
        body = m.group(1)
        vals = re.findall(r"[-+]?\d+\.\d*(?:[dDeE][-+]?\d+)?", body)
        arr = np.array([float(v.replace("d", "e").replace("D", "e")) for v in vals])
        assert arr.size == 4 * 4 * 72, (name, arr.size)
        out[name] = arr
    return out


def main():
    from oracle import harness
    harness.build_ref()
    vec = parse_fortran_vectors(os.path.join(REF, "fortran", "test_mod.F90"))
    np.savez(os.path.join(HERE, "fortran_golden.npz"), **vec)
    exe = os.path.join(ROOT, "oracle", "_ref", "pointers_only")
    txt = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    txt = "\n".join(l for l in txt.splitlines() if "total time" not in l) + "\n"
    open(os.path.join(HERE, "pointers_only_stdout.txt"), "w").write(txt)
    ro = harness.RefOracle(72)
    s = ro.init(3)
    out = {}
    for call in (1, 2):
        ro.run(s, 1, 1)
        for n in harness.MUTATED:
            a = s.arrays[n]
            if n in ("elem_state_dp3d", "elem_state_v", "elem_state_T"):
                a = a[:, int(s.ctl[3])]
            out[f"call{call}_{n}"] = a.copy()
    np.savez_compressed(os.path.join(HERE, "ref_outputs_E3.npz"), **out)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
