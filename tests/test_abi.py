"""The C-ABI library loads and exports every symbol include/gemotion_b200.h declares (no compute)."""
import os
import re

import pytest


def test_exports_match_header(gm):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "gemotion_b200.h")).read()
    declared = set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gm_desc", "gm_times"}
    assert declared == set(gm.capi.EXPORTS), declared ^ set(gm.capi.EXPORTS)
    if not os.path.exists(gm.capi.LIB_PATH):
        gm.capi.build()
    L = gm.capi.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.gm_version() == 110


def test_desc_struct_size(gm):
    import ctypes as C
    # layout agreed with the C side (gm_create rejects a mismatching struct_size)
    assert C.sizeof(gm.capi.gm_desc) % 8 == 0


def test_no_cpu_fallback(gm):
    """Without a GPU the product path must fail loudly, never fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cases
    sp = cases.cart(3)
    with pytest.raises(gm.capi.GmError):
        gm.operator.B200FEOperator(sp)
