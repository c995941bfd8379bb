"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol that
include/tandem_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tandem_b200._lib import LIB_PATH, declared_symbols, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    l = lib()
    assert l._tdm_missing == []
    raw = ctypes.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/tandem_b200.h but not exported"
    assert b"sm_100a" in l.tdm_version()


def test_header_cites_reference_interfaces():
    txt = open(os.path.join(ROOT, "include", "tandem_b200.h")).read()
    for cite in ("dr_mvsnet.h:36-66", "dr_fusion.h:44-73", "cuda_coarse_tracker.h:9-82"):
        assert cite in txt


def test_no_cpu_fallback():
    from tandem_b200 import CudaCoarseTracker, DrFusion, DrFusionOptions, DrMvsnet, TandemError, default_weights
    if lib().tdm_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(TandemError, match="no CPU fallback"):
        DrMvsnet(default_weights())
    with pytest.raises(TandemError, match="no CPU fallback"):
        DrFusion(DrFusionOptions())
    t = CudaCoarseTracker(640, 480)
    with pytest.raises(TandemError, match="no CPU fallback"):
        t.init()


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pat = re.compile(r"^\s*(from|import)\s+oracle|oracle/|oracle\.", re.M)
    for dp, _, fs in os.walk(os.path.join(ROOT, "tandem_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                code = "\n".join(l for l in src.splitlines() if not l.strip().startswith(("//", "#", "*", '"""')))
                assert not re.search(r"^\s*(from|import)\s+oracle", code, re.M), f"{f} imports the oracle"
                assert "oracle/_build" not in code and "libtsdf_oracle" not in code, f
