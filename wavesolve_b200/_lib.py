"""ctypes binding of ``libwavesolve_b200.so`` (the C ABI declared in
``include/wavesolve_b200.h``).  There is no CPU fallback: if the library is missing or a call
fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libwavesolve_b200.so")

_lib = None

ASM_TRANSPOSE = 1
ASM_FAST = 2

c_i32p = C.c_void_p   # device pointers travel as integers
c_f64p = C.c_void_p
c_i64 = C.c_int64
c_int = C.c_int
c_vp = C.c_void_p

class EigJob(C.Structure):
    """``ws_eig_job`` of include/wavesolve_b200.h (one eigenproblem of ``ws_eig_solve_batch``)."""
    _fields_ = [("pattern", C.c_void_p), ("solver", C.c_void_p), ("B_vals", C.c_void_p), ("sigma", C.c_double),
                ("kind", C.c_int), ("nev", C.c_int), ("ncv", C.c_int), ("maxit", C.c_int), ("tol", C.c_double),
                ("seed", C.c_uint64), ("edges", C.c_void_p), ("xy", C.c_void_p), ("Ntt", C.c_int64),
                ("w", C.c_void_p), ("V", C.c_void_p), ("stream", C.c_void_p), ("resid", C.c_void_p),
                ("workspace", C.c_void_p), ("info", C.c_int64 * 5), ("rc", C.c_int), ("error", C.c_char * 228)]


_SIGNATURES = {
    "ws_version": (c_int, []),
    "ws_last_error": (C.c_char_p, []),
    "ws_launch_count": (C.c_longlong, []),
    "ws_edges_first_appearance": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, C.POINTER(c_i64), c_int, c_vp]),
    "ws_pattern_create": (c_int, [c_vp, c_i64, c_i64, C.POINTER(c_vp), c_int, c_vp]),
    "ws_pattern_query": (c_int, [c_vp, C.POINTER(c_i64), C.POINTER(c_i64), C.POINTER(c_i64)]),
    "ws_pattern_export": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "ws_pattern_arrays": (c_int, [c_vp, C.POINTER(c_vp), C.POINTER(c_vp)]),
    "ws_pattern_destroy": (c_int, [c_vp]),
    "ws_assemble_scalar_p2": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "ws_assemble_scalar_p1": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "ws_assemble_mass_p2": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "ws_assemble_vector_p1": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_vp]),
    "ws_prune_zeros": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(c_i64), c_int, c_vp]),
    "ws_axpby": (c_int, [c_i64, C.c_double, c_vp, C.c_double, c_vp, c_int, c_vp]),
    "ws_spmv": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "ws_symbolic_create": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_i64, c_i64, c_int, C.POINTER(c_vp)]),
    "ws_symbolic_create_device": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_i64, c_i64, c_int, C.POINTER(c_vp),
                                          c_int, c_vp]),
    "ws_symbolic_query": (c_int, [c_vp, c_vp, c_vp]),
    "ws_symbolic_export": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ws_symbolic_destroy": (c_int, [c_vp]),
    "ws_solver_create": (c_int, [c_vp, c_vp, C.POINTER(c_vp), c_int, c_vp]),
    "ws_solver_factor": (c_int, [c_vp, c_vp, c_vp, C.c_double, c_int, c_vp]),
    "ws_solver_solve": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "ws_solver_check": (c_int, [c_vp, c_int, c_vp]),
    "ws_solver_stats": (c_int, [c_vp, c_vp, c_vp]),
    "ws_solver_quality": (c_int, [c_vp, c_vp, c_vp]),
    "ws_solver_set_overlap": (c_int, [c_vp, c_int]),
    "ws_solver_destroy": (c_int, [c_vp]),
    "ws_assemble_vector_qd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, C.c_double, c_vp, c_vp, c_int, c_int, c_vp]),
    "ws_eig_solve": (c_int, [c_vp, c_vp, c_vp, C.c_double, c_int, c_int, c_int, C.c_double, c_int, C.c_uint64,
                             c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "ws_host_eig_general": (c_int, [c_int, c_vp, c_vp, c_vp, c_vp, c_int]),
    "ws_eig_solve_batch": (c_int, [c_vp, c_int, c_int]),
    "ws_eig_workspace_create": (c_int, [C.POINTER(c_vp)]),
    "ws_eig_workspace_destroy": (c_int, [c_vp]),
    "ws_solver_factor_begin": (c_int, [c_vp, c_vp, c_vp, C.c_double, c_int, c_vp]),
    "ws_solver_factor_end": (c_int, [c_vp, c_int, c_vp]),
    "ws_locator_create": (c_int, [c_vp, c_vp, c_i64, c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.POINTER(c_vp), c_int, c_vp]),
    "ws_locator_destroy": (c_int, [c_vp]),
    "ws_locate": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_vp, c_i64, c_i64, C.c_double, c_int, c_vp, c_int, c_vp]),
    "ws_interp_weights": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_int, c_vp]),
    "ws_interp_apply": (c_int, [c_vp, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_int, c_vp]),
    "ws_overlap": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_int, c_vp]),
    "ws_bnormalize": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_int, c_int, c_vp]),
    "ws_vector_T": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_int, c_vp]),
    "ws_copy_to_host": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
}


class WavesolveB200Error(RuntimeError):
    """A C-ABI call failed.  ``rc`` is the library's status code (include/wavesolve_b200.h: 1 CUDA, 2 numeric,
    3 no convergence, < 0 argument / state).  For rc == 3 (``ws_eig_solve``: fewer than ``nev`` pairs converged)
    the output arrays have still been written; the drivers attach them as ``eigenvalues`` / ``eigenvectors``
    (the names of scipy's ``ArpackNoConvergence``, which the reference lets propagate from eigsh / eigs)."""
    rc = 0
    eigenvalues = None
    eigenvectors = None
    info = None


ERR_ARG, ERR_STATE, ERR_CUDA, ERR_NUMERIC, ERR_NOCONV = -1, -2, 1, 2, 3


def lib():
    """Load the library once.  Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WavesolveB200Error(
                "%s is missing: build it with `python -m wavesolve_b200._build` "
                "(or __graft_entry__.build()); wavesolve_b200 has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)     # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ws_last_error()
        e = WavesolveB200Error("wavesolve_b200 C-ABI call failed (%d): %s"
                               % (rc, msg.decode(errors="replace") if msg else "?"))
        e.rc = int(rc)
        raise e


def launch_count() -> int:
    return int(lib().ws_launch_count())
