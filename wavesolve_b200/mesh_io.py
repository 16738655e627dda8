"""Mesh hand-off: the regrouping of ``load_meshio_mesh`` (reference
waveguide.py:91-126) and a compact ``.npz`` container for the arrays the hot path reads.

Host side, numpy only -- meshing and file I/O stay on the host in the reference too.  The point of the
``.npz`` form is the marshalling cost per problem: a mesh object coming from pygmsh carries int64 / uint64
connectivity, one Python index list per material and no edge table, and every solve re-derives the
int32 arrays, the material-major order and (vector solver) the edge table from it.  The container stores
exactly what ``ScalarProblem`` / ``VectorProblem`` upload: points, int32 connectivity, one small material
id per element (or the explicit index lists when the sets are not a sorted partition) and, when present,
the edge table of ``get_unique_edges`` -- so a sweep worker goes from file to device arrays without
touching Python lists.
"""
from __future__ import annotations

import numpy as np

from .meshgen import CellBlock, Mesh

FORMAT_VERSION = 1


# ---------------------------------------------------------------------------------------
# load_meshio_mesh (waveguide.py:91-126)
# ---------------------------------------------------------------------------------------
def regroup_cell_sets(mesh):
    """The in-memory part of ``load_meshio_mesh`` (waveguide.py:93-126), applied to an object read by
    ``meshio.read`` from a gmsh ``.msh`` file: ``cell_sets[label]`` there holds one index array PER CELL
    BLOCK and the last set lists gmsh's bounding entities.  Afterwards the mesh looks like a pygmsh one:
    all triangles of all labelled blocks stacked in ``cells[1].data`` (block by block, label by label),
    ``cell_sets[label] = [None, rows, None]``, every other cell block replaced by ``None``.
    Mutates and returns ``mesh``."""
    keys = list(mesh.cell_sets.keys())[:-1]           # the last set: bounding entities (waveguide.py:93)
    chunks, rows = [], {key: [] for key in keys}
    total = 0
    for i in range(len(mesh.cells)):
        for key in keys:
            idx = mesh.cell_sets[key][i]
            n = len(idx)
            if n == 0:
                continue
            chunks.append(mesh.cells[i].data[idx])
            rows[key].append(np.arange(total, total + n))
            total += n
    data = (chunks[0] if len(chunks) == 1 else np.concatenate(chunks)) if chunks else []
    sets = {}
    for key in keys:
        r = rows[key]
        sets[key] = [None, (r[0] if len(r) == 1 else np.concatenate(r)) if r else [], None]
    mesh.cell_sets = sets
    mesh.cells[1].data = data
    for i in range(len(mesh.cells)):
        if i != 1:
            mesh.cells[i] = None
    return mesh


def load_meshio_mesh(meshname):
    """waveguide.py:91-126: read ``meshname + '.msh'`` with meshio and regroup it."""
    try:
        import meshio
    except ImportError as e:          # same failure the reference has at import time (waveguide.py:4)
        raise ImportError("load_meshio_mesh needs the meshio package") from e
    return regroup_cell_sets(meshio.read(meshname + ".msh"))


# ---------------------------------------------------------------------------------------
# compact container
# ---------------------------------------------------------------------------------------
def _small_int(a, signed=False):
    a = np.asarray(a)
    if a.size == 0:
        return a.astype(np.int32)
    hi = int(a.max())
    if int(a.min()) < 0 or hi >= 2 ** 31:
        if int(a.min()) < 0 and not signed:
            raise ValueError("negative index in a mesh array")
        return a.astype(np.int64)
    return a.astype(np.int32)


def mesh_to_arrays(mesh):
    """dict of numpy arrays describing ``mesh`` (what :func:`save_mesh_npz` writes)."""
    tris = np.asarray(mesh.cells[1].data)
    ne = len(tris)
    out = {"format_version": np.array(FORMAT_VERSION), "points": np.ascontiguousarray(mesh.points, dtype=np.float64),
           "tris": _small_int(tris), "labels": np.array([str(k) for k in mesh.cell_sets.keys()])}
    sets = []
    for key in mesh.cell_sets.keys():
        idx = mesh.cell_sets[key][1]
        sets.append(np.arange(ne, dtype=np.int64) if idx is None else np.asarray(idx, dtype=np.int64).ravel())
    # one id per element when the sets are a partition listed in ascending order (the usual case)
    total = sum(len(s) for s in sets)
    partition = total == ne and all(len(s) < 2 or bool(np.all(s[1:] > s[:-1])) for s in sets)
    if partition and ne:
        mat = np.full(ne, -1, dtype=np.int64)
        for i, s in enumerate(sets):
            if len(s) and (int(s[0]) < 0 or int(s[-1]) >= ne):
                partition = False
                break
            mat[s] = i
        partition = partition and bool(np.all(mat >= 0))
    if partition and len(sets) <= 255:
        out["material_id"] = (mat if ne else np.zeros(0, dtype=np.int64)).astype(np.uint8)
    else:
        for i, s in enumerate(sets):
            out["set_%d" % i] = _small_int(s, signed=True)
    edges = mesh.cells[0].data if mesh.cells[0] is not None else None
    if getattr(mesh, "edge_indices", None) is not None and edges is not None and len(np.asarray(edges)):
        out["edges"] = _small_int(edges)
        out["edge_indices"] = _small_int(mesh.edge_indices)
        if getattr(mesh, "edge_flips", None) is not None:
            out["edge_flips"] = np.asarray(mesh.edge_flips).astype(np.int8)
    return out


def mesh_from_arrays(z):
    """Inverse of :func:`mesh_to_arrays` (``z``: dict or ``NpzFile``): a mesh object with the attributes the
    hot path reads, including the edge table when it was stored."""
    files = set(z.files) if hasattr(z, "files") else set(z.keys())
    if int(z["format_version"]) != FORMAT_VERSION:
        raise ValueError("unknown mesh container version %s" % z["format_version"])
    labels = [str(s) for s in z["labels"]]
    if "material_id" in files:
        mat = np.asarray(z["material_id"])
        sets = {lab: [None, np.nonzero(mat == i)[0], None] for i, lab in enumerate(labels)}
    else:
        sets = {lab: [None, np.asarray(z["set_%d" % i]).astype(np.int64), None] for i, lab in enumerate(labels)}
    mesh = Mesh(np.asarray(z["points"]), np.asarray(z["tris"]), sets)
    if "edges" in files:
        mesh.cells[0] = CellBlock("line", np.asarray(z["edges"]))
        mesh.edge_indices = np.asarray(z["edge_indices"])
        if "edge_flips" in files:
            mesh.edge_flips = np.asarray(z["edge_flips"])
        mesh.num_edges = len(mesh.cells[0].data)        # waveguide.py:41-42
        mesh.num_points = mesh.points.shape[0]
    return mesh


def save_mesh_npz(path, mesh, compressed=False):
    """Write ``mesh`` (after ``get_unique_edges`` if the vector solver will use it) to ``path``."""
    (np.savez_compressed if compressed else np.savez)(path, **mesh_to_arrays(mesh))


def load_mesh_npz(path):
    with np.load(path, allow_pickle=False) as z:
        return mesh_from_arrays(z)
