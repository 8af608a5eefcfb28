"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol
include/caar_b200.h declares, and fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tinman_sandbox_b200 as tb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def test_header_symbols_match_binding_and_library():
    hdr = open(os.path.join(ROOT, "include", "caar_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(caar_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(tb.EXPORTED_SYMBOLS), declared ^ set(tb.EXPORTED_SYMBOLS)
    lib = tb.load_library()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.caar_version()


def test_field_counts_match_reference_extents():
    lib = tb.load_library()
    from tinman_sandbox_b200.capi import Dims
    d = Dims(7, 72, 4, 1, 3)
    for i, n in enumerate(tb.FIELD_NAMES):
        assert lib.caar_field_count(C.byref(d), i) == int(np.prod(tb.field_shape(n, 7, 72)))
    assert lib.caar_field_count(C.byref(d), 16) == 0
    # footprint per element quoted in SURVEY §8a: 186,112 B at nlev=72
    d1 = Dims(1, 72, 4, 1, 3)
    assert sum(lib.caar_field_count(C.byref(d1), i) for i in range(16)) * 8 == 186112


def test_no_cpu_fallback_without_device():
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(tb.CaarError) as ei:
        tb.Caar(2)
    assert "cuda" in str(ei.value).lower() or "device" in str(ei.value).lower()
    x = np.ones(8)
    with pytest.raises(tb.CaarError):
        tb.saxpby_host(3.0, 5.0, x, x.copy())


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under tinman_sandbox_b200/ or include/ may reference it."""
    for base in ("tinman_sandbox_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f)).read()
                    assert "oracle/" not in txt and "from oracle" not in txt and "import oracle" not in txt, (dp, f)


def test_library_is_built_for_sm_100a_with_tma_clusters_and_dmma():
    """The built library is what DESIGN.md says it is: sm_100a SASS only, the fused kernel instances move their tiles with
    TMA tensor copies (UTMALDG / UTMASTG), prefetch into L2 (UBLKPF), exchange scan totals between the CTAs of a cluster
    with st.async (STAS) and differentiate along igp on the FP64 tensor core (DMMA) — tools/sass_census.py, the table in
    profiles/r2v_sass_census.txt. A build that silently lost one of these (a flag, an #if) fails here, on the CPU."""
    import shutil
    import sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_census
    arch, per = sass_census.census(tb.lib_path())
    assert arch == ["sm_100a"], arch
    fused = per["caar_fused_kernel"]
    assert fused["functions"] == 28                      # 14 level counts x {Lagrangian, Eulerian}
    for op in ("UTMALDG", "UTMASTG", "UBLKPF", "STAS", "SYNCS", "DMMA"):
        assert fused[op] > 0, op
    assert fused["DMMA"] >= 20 * fused["functions"]      # 5 igp derivatives x 4 columns per instance
    for fam in ("levelop_kernel", "laplace_flat_kernel"):
        assert per[fam]["UTMALDG"] > 0 and per[fam]["UTMASTG"] > 0, fam
    assert per["caar_strict_kernel"]["DMMA"] == 0        # the bit-exact anchor stays on plain FP64 arithmetic
