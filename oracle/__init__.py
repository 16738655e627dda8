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
make_golden_at_size.py  one-off CPU runs of the BASELINE configs at their STATED sizes (C2 50k and C4 1M triangles
              through restated.py, C3 200k through the literal solve_waveguide): eigenvalues, mode count, mesh
              digests and the CPU timing (1545 s for C4 -- the same-configuration CPU baseline) ->
              tests/golden/c2_vec_50k.npz, c3_lantern_200k.npz, c4_vec_1M.npz.

Parity status: PINNED -- restated.py is checked against fixtures the literal reference produced
(tests/test_oracle_golden.py), against the literal reference run live on random meshes in the build container
(tests/test_oracle_live_reference.py), and its at-size eigenvalues are what the GPU tests assert.
"""
