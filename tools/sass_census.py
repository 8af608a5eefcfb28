#!/usr/bin/env python
"""Counts the SASS mnemonics that prove what the built library runs on (B200_PROFILING.md: UTMALDG/UTMASTG = TMA tensor
copies, UBLKPF = bulk L2 prefetch, STAS = st.async into a peer CTA's shared memory, UCGABAR = cluster barrier,
SYNCS = mbarrier, DMMA = FP64 tensor-core MMA) per kernel family of tinman_sandbox_b200/libcaar_b200.so.

    python tools/sass_census.py [path/to/lib.so]  > profiles/rN_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("UTMALDG", "UTMASTG", "UBLKPF", "STAS", "UCGABAR", "SYNCS", "DMMA", "DFMA", "SHFL", "LDS", "STS", "LDG", "STG",
        "STL", "LDL")


def census(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    fam, arch = None, set()
    for line in out.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            k = re.search(r"(\w+)(?:<[^(]*>)?\(", name.replace("(anonymous namespace)::", ""))
            fam = k.group(1) if k else name[:40]
            per.setdefault(fam, collections.Counter())["functions"] += 1
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and fam:
            op = m.group(1)
            for w in WANT:
                if op == w or op.startswith(w + ".") or op.startswith(w + "_"):
                    per[fam][w] += 1
    return sorted(arch), per


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tinman_sandbox_b200", "libcaar_b200.so")
    arch, per = census(lib)
    print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic census (cuobjdump -sass), arch = {', '.join(arch)}")
    print("%-26s %5s " % ("kernel family", "fns") + " ".join("%8s" % w for w in WANT))
    for fam, c in per.items():
        print("%-26s %5d " % (fam, c["functions"]) + " ".join("%8d" % c[w] for w in WANT))


if __name__ == "__main__":
    main()
