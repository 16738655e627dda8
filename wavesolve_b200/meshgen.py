"""Synthetic waveguide cross-section meshes (host side).

gmsh/pygmsh are not available in this environment, so the benchmark and the tests build
their meshes here.  The objects returned follow the mesh contract the reference's hot path
reads (documented by use in reference fe_solver.py:31-36, 156-160, 179-180 and by
waveguide.py:91-126 ``load_meshio_mesh``):

* ``mesh.points``           (Np, 3) float64, only ``[:, :2]`` is used
* ``mesh.cells[1].data``    (Ne, 6) or (Ne, 3) integer connectivity, gmsh node order
                            v0, v1, v2, m01, m12, m20, counter-clockwise
* ``mesh.cell_sets``        dict label -> ``[None, idx, None]`` (insertion order = material
                            processing order), ``idx`` indexes rows of ``cells[1].data``
* ``mesh.cells[0].data`` / ``mesh.edge_indices`` / ``num_edges`` / ``num_points`` are filled
  by ``get_unique_edges`` (reference waveguide.py:10-43) for the vector solver.

Meshing stays on the host in the reference as well (north_star), so this module is plain
numpy/scipy and is not part of the accelerated path.
"""
from __future__ import annotations

import numpy as np


class CellBlock:
    """Minimal stand-in for ``meshio.CellBlock`` (only ``.type`` and ``.data`` are read)."""

    def __init__(self, type_, data):
        self.type = type_
        self.data = data

    def __repr__(self):
        return "<CellBlock %s %s>" % (self.type, getattr(self.data, "shape", None))


class Mesh:
    """Minimal stand-in for ``meshio.Mesh`` as produced by pygmsh and mutated by wavesolve."""

    def __init__(self, points, tris, cell_sets):
        self.points = points
        kind = "triangle6" if tris.shape[1] == 6 else "triangle"
        # pygmsh meshes carry [line, triangle, vertex] blocks; only index 1 is read by the
        # assembly, index 0 is overwritten by get_unique_edges.
        self.cells = [CellBlock("line", np.zeros((0, 2), dtype=np.int64)), CellBlock(kind, tris)]
        self.cell_sets = cell_sets

    @property
    def num_elements(self):
        return self.cells[1].data.shape[0]


def _ccw(points, tris):
    """Force counter-clockwise orientation (the reference assumes J > 0, shape_funcs.py:202)."""
    p = points[:, :2]
    a, b, c = p[tris[:, 0]], p[tris[:, 1]], p[tris[:, 2]]
    J = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    flip = J < 0
    tris = tris.copy()
    tris[flip, 1], tris[flip, 2] = tris[flip, 2].copy(), tris[flip, 1].copy()
    return tris[np.abs(J) > 0]


def _add_midside_nodes(points, tris3):
    """P1 -> P2: one straight-edge midside node per unique edge, numbered after the
    vertices in first-appearance order of the (sorted) edge; element node order
    v0,v1,v2,m01,m12,m20 (gmsh triangle6)."""
    ne = tris3.shape[0]
    ed = np.stack([tris3[:, [0, 1]], tris3[:, [1, 2]], tris3[:, [2, 0]]], axis=1).reshape(-1, 2)
    ed = np.sort(ed, axis=1)
    npts = points.shape[0]
    key = ed[:, 0].astype(np.int64) * npts + ed[:, 1]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    eid = rank[inv].reshape(ne, 3)
    ued = np.empty((len(uniq), 2), dtype=np.int64)
    ued[rank] = ed[first]
    mid = 0.5 * (points[ued[:, 0]] + points[ued[:, 1]])
    pts = np.vstack([points, mid])
    tris6 = np.hstack([tris3, eid + npts])
    return pts, tris6


def _spatial_sort(points, tris, bits=16):
    """Order the triangles along a Morton (Z-order) curve of their centroids.  Mesh generators
    emit elements in a spatially coherent order (advancing front / incremental Delaunay); qhull
    does not, and an incoherent element order makes the first-appearance edge numbering -- and
    every gather that follows -- needlessly cache-hostile."""
    cen = points[tris[:, :3], :2].mean(axis=1)
    lo = cen.min(axis=0)
    span = np.maximum(cen.max(axis=0) - lo, 1e-300)
    q = np.minimum(((cen - lo) / span * (1 << bits)).astype(np.uint64), (1 << bits) - 1)

    def spread(v):
        v = v & np.uint64(0xFFFFFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v

    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1))
    return tris[np.argsort(key, kind="stable")]


def _materials_by_radius(points, tris, radii_labels, centres=None):
    """Assign each triangle to the first (label, radius[, centre]) disc containing its centroid.
    Returns an ordered dict label -> [None, idx, None]."""
    cen = points[tris[:, :3], :2].mean(axis=1)
    taken = np.zeros(len(tris), dtype=bool)
    sets = {}
    for item in radii_labels:
        label, rad = item[0], item[1]
        ctrs = item[2] if len(item) > 2 else [(0.0, 0.0)]
        inside = np.zeros(len(tris), dtype=bool)
        for (cx, cy) in ctrs:
            inside |= (cen[:, 0] - cx) ** 2 + (cen[:, 1] - cy) ** 2 < rad * rad
        sel = inside & ~taken
        taken |= sel
        prev = sets.get(label)
        idx = np.nonzero(sel)[0]
        if prev is not None:
            idx = np.concatenate([prev[1], idx])
        sets[label] = [None, idx, None]
    return sets


def structured_square(n, half_width=10.0, r_core=5.0, order=2, dtype=np.int64):
    """Right-triangle grid on [-hw, hw]^2 with a core disc: 2 n^2 triangles.
    Maximises exactly-cancelling entries (stress test for the zero-pruning rules of the
    vector assembly, reference fe_solver.py:170-213)."""
    xs = np.linspace(-half_width, half_width, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    pts = np.zeros(((n + 1) ** 2, 3))
    pts[:, 0] = X.ravel()
    pts[:, 1] = Y.ravel()
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (n + 1)
    v11 = v01 + 1
    t1 = np.stack([v00, v10, v11], axis=1)
    t2 = np.stack([v00, v11, v01], axis=1)
    tris = np.empty((2 * n * n, 3), dtype=np.int64)
    tris[0::2] = t1
    tris[1::2] = t2
    tris = _ccw(pts, tris)
    if order == 2:
        pts, tris = _add_midside_nodes(pts, tris)
    sets = _materials_by_radius(pts, tris, [("core", r_core), ("cladding", 1e300)])
    return Mesh(pts, tris.astype(dtype), sets)


def _disc_points(h, R, rng, jitter, circles):
    """Jittered square-grid points inside radius R plus equispaced points on each circle in
    ``circles`` = [(cx, cy, r)], with grid points closer than 0.55 h to a circle removed."""
    nx = int(np.ceil(R / h)) + 1
    g = np.arange(-nx, nx + 1) * h
    X, Y = np.meshgrid(g, g, indexing="xy")
    # offset alternate rows (hex-like packing gives well shaped Delaunay triangles)
    X = X + 0.5 * h * (np.arange(X.shape[0])[:, None] % 2)
    P = np.stack([X.ravel(), Y.ravel()], axis=1)
    P += (rng.random(P.shape) - 0.5) * (2 * jitter * h)
    keep = P[:, 0] ** 2 + P[:, 1] ** 2 < (R - 0.55 * h) ** 2
    onc = []
    for (cx, cy, r) in circles:
        d = np.abs(np.hypot(P[:, 0] - cx, P[:, 1] - cy) - r)
        keep &= d > 0.55 * h
        m = max(8, int(np.ceil(2 * np.pi * r / h)))
        th = 2 * np.pi * np.arange(m) / m
        onc.append(np.stack([cx + r * np.cos(th), cy + r * np.sin(th)], axis=1))
    return np.vstack([P[keep]] + onc)


def fiber_disc(target_elements, r_core=5.0, r_clad=10.0, order=2, seed=0, jitter=0.2,
               dtype=np.int64, spatial_sort=False):
    """Step-index circular fibre (getting-started notebook cell 9: r_core=5, r_clad=10):
    Delaunay triangulation of jittered points in the cladding disc, conforming (up to
    Delaunay) to the core circle.  ``target_elements`` is approximate."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    npts = max(16, target_elements // 2)
    h = np.sqrt(np.pi * r_clad ** 2 / npts)
    P = _disc_points(h, r_clad, rng, jitter, [(0.0, 0.0, r_clad), (0.0, 0.0, r_core)])
    tri = Delaunay(P)
    pts = np.zeros((P.shape[0], 3))
    pts[:, :2] = P
    tris = _ccw(pts, tri.simplices.astype(np.int64))
    # drop slivers on the hull (nearly collinear boundary points)
    p = pts[:, :2]
    a, b, c = p[tris[:, 0]], p[tris[:, 1]], p[tris[:, 2]]
    J = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    tris = tris[J > 1e-3 * h * h]
    if spatial_sort:
        tris = _spatial_sort(pts, tris)
    if order == 2:
        pts, tris = _add_midside_nodes(pts, tris)
    sets = _materials_by_radius(pts, tris, [("core", r_core), ("cladding", 1e300)])
    return Mesh(pts, tris.astype(dtype), sets)


def lantern_cross_section(target_elements, n_cores=7, order=2, seed=0, jitter=0.2,
                          r_jacket=40.0, r_clad_bundle=13.0, r_core=1.06, dtype=np.int64):
    """Photonic-lantern cross-section in the spirit of lantern-demo notebook cell 1
    (``FiberBundleLantern``, reference waveguide.py:728-837): a low-index jacket disc, a
    cladding region formed by the union of ``n_cores`` discs (centre + ring), and one small
    core per disc.  Geometry is scaled so the bundle radius is ``r_clad_bundle``.
    Labels: 'core', 'cladding', 'jacket'."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    ring = n_cores - 1
    rc = r_clad_bundle / 3.0              # radius of each fibre cladding disc
    centres = [(0.0, 0.0)]
    for i in range(ring):
        th = 2 * np.pi * i / ring
        centres.append((2 * rc * np.cos(th), 2 * rc * np.sin(th)))
    npts = max(64, target_elements // 2)
    h = np.sqrt(np.pi * r_jacket ** 2 / npts)
    # graded density would be nicer; a uniform mesh keeps the generator simple and the
    # element count predictable.
    circles = [(0.0, 0.0, r_jacket)]
    circles += [(cx, cy, 1.25 * rc) for (cx, cy) in centres]
    circles += [(cx, cy, r_core) for (cx, cy) in centres if r_core > 1.5 * h]
    P = _disc_points(h, r_jacket, rng, jitter, circles)
    tri = Delaunay(P)
    pts = np.zeros((P.shape[0], 3))
    pts[:, :2] = P
    tris = _ccw(pts, tri.simplices.astype(np.int64))
    p = pts[:, :2]
    a, b, c = p[tris[:, 0]], p[tris[:, 1]], p[tris[:, 2]]
    J = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    tris = tris[J > 1e-3 * h * h]
    if order == 2:
        pts, tris = _add_midside_nodes(pts, tris)
    sets = _materials_by_radius(
        pts, tris,
        [("core", r_core, centres), ("cladding", 1.25 * rc, centres), ("jacket", 1e300)])
    return Mesh(pts, tris.astype(dtype), sets)


def holey_fiber(target_elements, pitch=3.0, r_hole=1.0, rings=2, r_clad=9.0, order=2, seed=0, jitter=0.2,
                hollow_core=False, dtype=np.int64):
    """Air-hole fibre (in the spirit of ``PhotonicCrystalFiber`` / ``PhotonicBandgapFiber``, reference
    waveguide.py:604-726 and getting-started.ipynb cells 15, 24): a silica disc with a hexagonal lattice of
    ``rings`` rings of air holes (n = 1.0) around the centre; the centre site is solid silica (index-guiding
    PCF) or, with ``hollow_core``, a larger air hole.  High index contrast and -- solved with ``target_neff``
    inside the band -- a shift-invert matrix that is strongly indefinite: the stress test for the
    factorisation's accuracy guard.  Labels: 'air', 'silica'."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    centres = []
    for i in range(-rings, rings + 1):
        for j in range(-rings, rings + 1):
            if max(abs(i), abs(j), abs(i + j)) <= rings and (i, j) != (0, 0):
                centres.append((pitch * (i + 0.5 * j), pitch * (np.sqrt(3.0) / 2.0) * j))
    hole_r = [(cx, cy, r_hole) for (cx, cy) in centres]
    if hollow_core:
        hole_r.append((0.0, 0.0, 1.4 * r_hole))
    npts = max(64, target_elements // 2)
    h = np.sqrt(np.pi * r_clad ** 2 / npts)
    P = _disc_points(h, r_clad, rng, jitter, [(0.0, 0.0, r_clad)] + hole_r)
    tri = Delaunay(P)
    pts = np.zeros((P.shape[0], 3))
    pts[:, :2] = P
    tris = _ccw(pts, tri.simplices.astype(np.int64))
    p = pts[:, :2]
    a, b, c = p[tris[:, 0]], p[tris[:, 1]], p[tris[:, 2]]
    J = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    tris = tris[J > 1e-3 * h * h]
    if order == 2:
        pts, tris = _add_midside_nodes(pts, tris)
    cen = pts[tris[:, :3], :2].mean(axis=1)
    air = np.zeros(len(tris), dtype=bool)
    for (cx, cy, r) in hole_r:
        air |= (cen[:, 0] - cx) ** 2 + (cen[:, 1] - cy) ** 2 < r * r
    sets = {"air": [None, np.nonzero(air)[0], None], "silica": [None, np.nonzero(~air)[0], None]}
    return Mesh(pts, tris.astype(dtype), sets)


FIBER_IOR = {"core": 1.444 + 8.8e-3, "cladding": 1.444}          # getting-started cell 9
HOLEY_IOR = {"air": 1.0, "silica": 1.444}                         # getting-started cells 15 / 24
LANTERN_IOR = {"core": 1.4521, "cladding": 1.44692, "jacket": 1.44692 - 5e-3}   # lantern-demo cell 1
