"""CPU tier: the C-ABI library loads and exports every symbol include/ssm_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    txt = (ROOT / "include" / "ssm_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ssm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ssm_b200
    from ssm_b200 import _lib
    assert _lib.LIB_PATH.exists(), "build the CUDA library first: python __graft_entry__.py build"
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = header_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    # the Python twin binds exactly the declared surface
    assert sorted(ssm_b200.exported_symbols()) == declared
    assert ssm_b200.lib().ssm_version() >= 100


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing in the shipped package may import, link or execute it."""
    pkg = ROOT / "semiseparablematrices.jl_b200"
    for f in list(pkg.rglob("*.py")):
        txt = f.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + [pkg / "Makefile"]:
        assert "oracle/" not in f.read_text() and "ssm_oracle" not in f.read_text(), f


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product must fail loudly (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ssm_b200
    with pytest.raises(ssm_b200.SSMError):
        ssm_b200.Context(0)
