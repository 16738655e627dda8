"""TEST INFRASTRUCTURE ONLY.  Import the literal reference implementation.

The reference (``/root/reference/src/wavesolve``) is pure Python but has five
top-level imports that are absent here and are not needed by the numeric hot path:
matplotlib(.pyplot/.colors), numexpr (only in dead code, shape_funcs.py:100-150),
pygmsh, gmsh and meshio (meshing).  We satisfy them with empty stub modules and import
the reference unmodified from where it lies.  Nothing is copied.

Only usable where ``/root/reference`` exists (the build container); the GPU box uses the
committed fixtures under ``tests/golden`` instead.
"""
import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("WAVESOLVE_REFERENCE_SRC", "/root/reference/src")

_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.tri",
          "numexpr", "pygmsh", "pygmsh.occ", "gmsh", "meshio"]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "wavesolve", "fe_solver.py"))


class _Anything(types.ModuleType):
    """Module stub: any attribute is another stub (enough for ``import x.y as z`` and
    annotations like ``pygmsh.occ.Geometry`` evaluated at def time)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed module %s called: not available in the oracle shim" % self.__name__)


def load():
    """Return (fe_solver, shape_funcs, waveguide) modules of the literal reference."""
    if not available():
        raise RuntimeError("literal reference not present at %s" % REFERENCE_SRC)
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    fe = importlib.import_module("wavesolve.fe_solver")
    sf = importlib.import_module("wavesolve.shape_funcs")
    wg = importlib.import_module("wavesolve.waveguide")
    return fe, sf, wg
