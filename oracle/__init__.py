"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the wavesolve FEM hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker / the baseline being timed.
The product path (``wavesolve_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Contents
--------
refshim.py    import the *literal* reference (``/root/reference/src``) behind stub
              modules for its non-numeric imports.  Only works in the build
              container (the reference tree does not travel to the GPU box).
restated.py   numpy/scipy restatement of the reference algorithms
              (fe_solver.py, shape_funcs.py, waveguide.py), each function citing the
              reference file:line it follows.  Pinned against the literal reference by
              ``oracle/make_golden.py`` -> ``tests/golden/*.npz``.
make_golden.py  regenerates the golden fixtures from the literal reference.
"""
