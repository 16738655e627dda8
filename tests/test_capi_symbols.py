"""CPU: the C-ABI library loads, exports every symbol the header declares, and the product
path refuses to run without a CUDA device (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "wavesolve_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from wavesolve_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 10
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "missing symbol " + n
        assert n in _lib._SIGNATURES, "no ctypes signature for " + n
    assert L.ws_version() >= 100


def test_argument_errors_do_not_need_a_gpu():
    from wavesolve_b200 import _lib
    L = _lib.lib()
    assert L.ws_pattern_query(None, None, None, None) < 0
    assert b"pattern" in L.ws_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda():
    import wavesolve_b200 as w
    from wavesolve_b200 import _lib
    mesh = w.meshgen.structured_square(3)
    with pytest.raises(_lib.WavesolveB200Error):
        w.construct_AB(mesh, w.meshgen.FIBER_IOR, 4.0, sparse=True)
    with pytest.raises(_lib.WavesolveB200Error):
        w.get_unique_edges(w.meshgen.structured_square(3, order=1))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wavesolve_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
