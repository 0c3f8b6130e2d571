"""The C-ABI library loads on a CPU-only box and exports every symbol include/ctan_sm100.h declares;
the ctypes binding covers exactly that set.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ctan_sm100.h")).read()
    return re.findall(r"^int (ctan_\w+)\(", txt, re.M)


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 25 and len(set(names)) == len(names)
    txt = open(os.path.join(ROOT, "include", "ctan_sm100.h")).read()
    assert txt.count("replaces:") == len(names)          # every entry point cites the reference code it replaces


def test_library_exports_every_declared_symbol():
    import ctan_b200  # noqa: F401
    from ctan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(dll, name), name
    assert set(_declared()) == set(_lib.exported_symbols())
    _lib.lib()                                           # argtypes resolve for every bound symbol


def test_product_fails_loudly_without_cuda():
    import torch
    import ctan_b200  # noqa: F401
    from ctan_b200._lib import CtanLibraryError
    from ctan_b200.models import LastNeighborLoader
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(CtanLibraryError):
        LastNeighborLoader(10, 2, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "non-dissipative-propagation-ctdgs_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read(), f
