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


def test_batch_api_structs_and_argument_errors():
    """ws_eig_job is mirrored field by field in _lib.EigJob (static_assert(sizeof == 400) on the C side); handles that
    own no device memory can be created and destroyed without a GPU; bad arguments come back as status codes."""
    from wavesolve_b200 import _lib
    L = _lib.lib()
    assert ctypes.sizeof(_lib.EigJob) == 400
    assert _lib.EigJob.info.offset == 128 and _lib.EigJob.rc.offset == 168 and _lib.EigJob.error.offset == 172
    h = ctypes.c_void_p()
    assert L.ws_eig_workspace_create(ctypes.byref(h)) == 0 and h.value
    assert L.ws_eig_workspace_destroy(h) == 0
    assert L.ws_eig_solve_batch(None, 0, 0) < 0
    assert L.ws_solver_quality(None, None, None) < 0 and b"solver" in L.ws_last_error()
    assert L.ws_solver_factor_end(None, 0, None) < 0
    assert L.ws_copy_to_host(None, None, -1, 0, 0, None) < 0


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
