"""Builds every native artefact of the package in-tree, for sm_100a only:

  tinman_sandbox_b200/libcaar_b200.so   CUDA kernels + C-ABI   (csrc/Makefile: nvcc -gencode
                                        arch=compute_100a,code=sm_100a -lineinfo)
  tinman_sandbox_b200/host/…            C++ host side (shim + driver), when host/Makefile exists

    python -m tinman_sandbox_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False, verbose=False):
    out = None if verbose else subprocess.DEVNULL
    csrc = os.path.join(HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=out)
    subprocess.check_call(["make", "-j4", "-C", csrc], stdout=out)
    host = os.path.join(HERE, "host")
    if os.path.exists(os.path.join(host, "Makefile")):
        if force:
            subprocess.check_call(["make", "-C", host, "clean"], stdout=out)
        subprocess.check_call(["make", "-C", host], stdout=out)
    return os.path.join(HERE, "libcaar_b200.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
