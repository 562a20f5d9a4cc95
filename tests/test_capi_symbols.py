"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for f in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", f)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from bitdance_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_status_strings_and_version(lib):
    assert lib.bd_abi_version() >= 1
    assert lib.bd_strerror(0).decode() == "ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert lib.bd_strerror(code).decode() not in ("ok", "unknown status")


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: ops on CPU tensors raise instead of computing."""
    import torch
    from bitdance_b200 import ops
    from bitdance_b200._lib import BitDanceNativeError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(BitDanceNativeError):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(64, 64, dtype=torch.bfloat16))
    from bitdance_b200.modeling.utils import MLPconnector
    with pytest.raises(RuntimeError):
        MLPconnector(32, 64, "gelu_pytorch_tanh")(torch.zeros(1, 4, 32))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under bitdance_b200/ or modeling/ may import it."""
    bad = []
    for base in ("bitdance_b200", "modeling"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
