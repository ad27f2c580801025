"""CPU: the C-ABI shared library loads without a GPU driver and exports every symbol that
include/atlas_b200.h declares.  No compute calls."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "atlas_b200.h")).read()
    return sorted(set(re.findall(r"\b(atlas_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    from atlas_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 8
    for sym in declared:
        assert hasattr(L, sym), f"{sym} declared in include/atlas_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared)


def test_workspace_size_is_pure_host():
    from atlas_b200 import _lib

    L = _lib.lib()
    small = L.atlas_b200_mips_workspace_bytes(10000, 64, 40)
    big = L.atlas_b200_mips_workspace_bytes(4 << 20, 256, 40)
    assert 0 < small < big < (1 << 31)
    assert L.atlas_b200_version().startswith(b"atlas_b200")


def test_no_cpu_fallback():
    import pytest
    import torch

    from atlas_b200 import ops
    from atlas_b200._lib import AtlasB200Error

    bank = torch.zeros(64, 768, dtype=torch.float16)
    with pytest.raises(AtlasB200Error):
        ops.mips_topk(bank, torch.zeros(1, 768), 4)
