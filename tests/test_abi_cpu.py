"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/galcamb.h declares,
the ctypes mirror of galcamb_model has the compiled size, and compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "galcamb.h")).read()
    syms = set(re.findall(r"\b(galcamb_[a-z_0-9]+_)\s*\(", h))
    # the reference's own symbols (camb/interface.f90) and the two table accessors, include/galileon_abi.h
    g = open(os.path.join(ROOT, "include", "galileon_abi.h")).read()
    syms |= set(re.findall(r"^(?:int|double|void)\s+([A-Za-z_]+_)\s*\(", g, flags=re.M))
    return sorted(syms)


def test_library_exports_every_declared_symbol(built_cuda_lib):
    L = C.CDLL(built_cuda_lib)
    syms = declared_symbols()
    assert len(syms) >= 18 + 15
    assert {"arrays_", "GetX_", "GetH_", "GetdHdX_", "ct_", "pigalprime_", "freegal_"} <= set(syms)
    for s in syms:
        assert hasattr(L, s), s


def test_struct_layout_matches_header(built_cuda_lib):
    from cosmomc_galileon_b200.camb_gpu import CModel, load_library
    assert load_library().galcamb_sizeof_model_() == C.sizeof(CModel)


def test_no_cpu_fallback(built_cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosmomc_galileon_b200 import GalCamb
    from cosmomc_galileon_b200.camb_gpu import GalCambError
    with pytest.raises((GalCambError, OSError)):
        GalCamb(0)


def test_fixture_roundtrip():
    from cosmomc_galileon_b200 import ModelTables
    t = ModelTables.load(os.path.join(ROOT, "tests", "golden", "config2_tables.npz"))
    c = t.cstruct()
    assert c.nk == len(t.d["k"]) and c.nt == len(t.d["tau"]) and c.nq == len(t.d["q"])
    assert c.use_galileon == 1 and c.ngal == 30000 and c.n_tau_regions >= 2
    assert np.all(np.diff(t.d["tau"]) > 0) and np.all(np.diff(t.d["k"]) > 0)


def test_product_does_not_import_oracle():
    # the product path must not route through oracle/: no source under the package mentions it
    pk = os.path.join(ROOT, "cosmomc_galileon_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in txt and "liborc" not in txt and "orc.hpp" not in txt, f
