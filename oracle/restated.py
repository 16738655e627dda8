"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of the reference hot path.

Each function cites the reference file:line it follows (paths relative to
``/root/reference/src/wavesolve``).  The restatement exists because (a) the literal
reference does not travel to the GPU box and (b) its vector path is O(E^2) / dense
(waveguide.py:27-32, fe_solver.py:206, fe_solver.py:398) and cannot run the 50k..1M element
configurations at all.  Parity of this file against the literal reference is pinned by
``oracle/make_golden.py`` -> ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py``.

The arithmetic *order* of the reference formulas is kept (same association, separate
multiply and subtract) so that element matrices come out bit-identical to the reference;
the summation of duplicate entries follows the reference's element order for the vector
path (sequential ``+=`` into lil matrices, fe_solver.py:196-199) so that the set of entries
that cancel to exactly 0.0 -- which the reference prunes -- is reproduced.

Pinned status (tests/test_oracle_golden.py): scalar P2 matrices bit-identical (pattern and
values) to the literal reference on every fixture; edge tables identical; vector matrices
identical in pattern (pruned zeros included), values identical except <1% of entries that
differ in the last bit because the literal code squares edge lengths with scalar ``l**2``
(libm pow, not correctly rounded) where this file multiplies; eigenvalues of both solve
paths agree to <=1e-10 relative in neff with the literal ``solve_waveguide`` /
``solve_waveguide_vec('transform')``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# --------------------------------------------------------------------------------------
# mesh helpers
# --------------------------------------------------------------------------------------
def material_major(mesh, attr_data=None):
    """Elements in the order the reference processes them: for each material in
    ``cell_sets`` insertion order, ``data[tuple(cell_sets[m])][0,:,0,:]``
    (fe_solver.py:36, fe_solver.py:179).  Returns (conn, per-element material index, labels)."""
    data = mesh.cells[1].data if attr_data is None else attr_data
    conns, mats, labels = [], [], list(mesh.cell_sets.keys())
    for mi, label in enumerate(labels):
        idx = mesh.cell_sets[label][1]
        c = data[idx]
        conns.append(c)
        mats.append(np.full(len(c), mi, dtype=np.int64))
    if not conns:
        return data[:0], np.zeros(0, dtype=np.int64), labels
    return np.vstack(conns), np.concatenate(mats), labels


# --------------------------------------------------------------------------------------
# edge table  (waveguide.py:10-43)
# --------------------------------------------------------------------------------------
def get_unique_edges(mesh, mutate=True):
    """Unique edges numbered by first appearance while scanning triangles in order and,
    inside a triangle, local edges (0,1),(1,2),(2,0); an edge is stored as a sorted pair
    (waveguide.py:20-32).  Vectorised: rank-by-first-index of the unique sorted pairs."""
    tris = np.asarray(mesh.cells[1].data)
    ne = tris.shape[0]
    loc = tris[:, [0, 1, 1, 2, 2, 0]].reshape(ne * 3, 2).astype(np.int64)
    srt = np.sort(loc, axis=1)
    base = int(srt.max()) + 1 if srt.size else 1
    key = srt[:, 0] * base + srt[:, 1]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq))
    edges = srt[first[order]].astype(tris.dtype if tris.size else np.int64)
    edge_indices = rank[inv].reshape(ne, 3).astype(np.uint)          # waveguide.py:16
    flips = np.where(loc[:, 0] <= loc[:, 1], 1.0, -1.0).reshape(ne, 3)  # waveguide.py:33-34
    if mutate:                                                        # waveguide.py:37-42
        mesh.cells[0].data = edges
        mesh.edge_indices = edge_indices
        mesh.edge_flips = flips
        mesh.num_edges = len(edges)
        mesh.num_points = mesh.points.shape[0]
    return edges, edge_indices


# --------------------------------------------------------------------------------------
# P2 scalar element matrices  (shape_funcs.py:192-246)
# --------------------------------------------------------------------------------------
def p2_element_matrices(tp):
    """tp: (ne, 6, >=2) coordinates gathered per element.  Returns NN, dNdN (ne,6,6).
    Straight-sided triangles: only the three vertices enter (shape_funcs.py:196-202)."""
    x1, y1 = tp[:, 0, 0], tp[:, 0, 1]
    x2, y2 = tp[:, 1, 0], tp[:, 1, 1]
    x3, y3 = tp[:, 2, 0], tp[:, 2, 1]
    x21, y21 = x2 - x1, y2 - y1
    x31, y31 = x3 - x1, y3 - y1
    x32, y32 = x3 - x2, y3 - y2
    J = x21 * y31 - x31 * y21
    ne = len(J)
    M = np.zeros((ne, 6, 6))
    K = np.zeros((ne, 6, 6))

    def sym(mat, pairs, val):
        for (i, j) in pairs:
            mat[:, i, j] = val
            mat[:, j, i] = val

    # mass (shape_funcs.py:205-214)
    sym(M, [(0, 0), (1, 1), (2, 2)], J / 60)
    sym(M, [(3, 3), (4, 4), (5, 5)], 4 * J / 45)
    sym(M, [(0, 1), (1, 2), (0, 2)], -J / 360)
    sym(M, [(0, 4), (1, 5), (2, 3)], -J / 90)
    sym(M, [(3, 4), (3, 5), (4, 5)], 2 * J / 45)
    # stiffness (shape_funcs.py:222-244); association kept as in the reference
    K[:, 0, 0] = (y32 * y32 + x32 * x32) / (2 * J)
    K[:, 1, 1] = (y31 * y31 + x31 * x31) / (2 * J)
    K[:, 2, 2] = (y21 * y21 + x21 * x21) / (2 * J)
    f = 4 / (3 * J)
    K[:, 3, 3] = f * (y32 ** 2 + y31 * y21 + x32 ** 2 + x31 * x21)
    K[:, 4, 4] = f * (y31 ** 2 - y21 * y32 + x31 ** 2 - x21 * x32)
    K[:, 5, 5] = f * (y21 ** 2 + y32 * y31 + x21 ** 2 + x32 * x31)
    sym(K, [(0, 1)], (y32 * y31 + x32 * x31) / (6 * J))
    sym(K, [(1, 2)], (y31 * y21 + x31 * x21) / (6 * J))
    sym(K, [(0, 2)], (-y21 * y32 - x21 * x32) / (6 * J))
    sym(K, [(0, 3), (1, 3)], -2 * (y32 * y31 + x32 * x31) / (3 * J))
    sym(K, [(0, 5), (2, 5)], 2 * (y21 * y32 + x21 * x32) / (3 * J))
    sym(K, [(2, 4), (1, 4)], -2 * (y31 * y21 + x31 * x21) / (3 * J))
    sym(K, [(3, 4)], f * (y21 * y32 + x21 * x32))
    sym(K, [(4, 5)], -f * (y31 * y32 + x31 * x32))
    sym(K, [(3, 5)], -f * (y31 * y21 + x31 * x21))
    return M, K


def construct_AB_order2(mesh, IOR_dict, k):
    """Scalar P2 assembly, sparse=True branch (fe_solver.py:20-69): material-major COO
    triplets, entry p of element e at row t[p%6], col t[p//6] (fe_solver.py:38-40);
    A_e = (k^2 n^2) NN - dNdN, B_e = NN (fe_solver.py:51-54); csr_matrix from COO with the
    shape inferred and duplicates summed, explicit zeros kept (fe_solver.py:67-68)."""
    conn, mat, labels = material_major(mesh)
    pts = mesh.points
    rows = np.repeat(conn, 6, 0).ravel()
    cols = np.repeat(conn, 6, 1).ravel()
    NN, KK = p2_element_matrices(pts[conn])
    A = np.empty_like(NN)
    for mi, label in enumerate(labels):
        sel = mat == mi
        A[sel] = (k ** 2 * IOR_dict[label] ** 2) * NN[sel] - KK[sel]
    Am = sp.csr_matrix((A.ravel(), (rows, cols)))
    Bm = sp.csr_matrix((NN.ravel(), (rows, cols)))
    return Am, Bm


def construct_B_order2(mesh):
    """Mass matrix over all cells, ignoring materials (fe_solver.py:71-99, sparse=True)."""
    conn = np.asarray(mesh.cells[1].data)
    NN, _ = p2_element_matrices(mesh.points[conn])
    rows = np.repeat(conn, 6, 0).ravel()
    cols = np.repeat(conn, 6, 1).ravel()
    return sp.csr_matrix((NN.ravel(), (rows, cols)))


# --------------------------------------------------------------------------------------
# P1 + Whitney vector element blocks  (shape_funcs.py:310-428)
# --------------------------------------------------------------------------------------
def vec_element_blocks(tp, conn):
    """tp: (ne,3,>=2) vertex coordinates, conn: (ne,3) global vertex ids.
    Returns NeNe, curlcurl, NedN, LNN, LdNdN, each (ne,3,3)."""
    x21 = tp[:, 1, 0] - tp[:, 0, 0]
    y21 = tp[:, 1, 1] - tp[:, 0, 1]
    x31 = tp[:, 2, 0] - tp[:, 0, 0]
    y31 = tp[:, 2, 1] - tp[:, 0, 1]
    x32 = tp[:, 2, 0] - tp[:, 1, 0]
    y32 = tp[:, 2, 1] - tp[:, 1, 1]
    c = np.asarray(conn).astype(np.int64)
    s1 = np.where(c[:, 0] < c[:, 1], 1.0, -1.0)          # shape_funcs.py:317-319
    s2 = np.where(c[:, 1] < c[:, 2], 1.0, -1.0)
    s3 = np.where(c[:, 2] < c[:, 0], 1.0, -1.0)
    l12 = np.sqrt(x21 * x21 + y21 * y21) * s1             # shape_funcs.py:320-322
    l23 = np.sqrt(x32 * x32 + y32 * y32) * s2
    l31 = np.sqrt(x31 * x31 + y31 * y31) * s3
    J = x21 * y31 - x31 * y21
    ne = len(J)
    sq12, sq23, sq31 = l12 * l12, l23 * l23, l31 * l31    # l**2 in the reference
    xx, yy = x21 * x31, y21 * y31
    x3, y3 = (3 * x21) * x31, (3 * y21) * y31      # python evaluates 3*x21*x31 left to right

    E = np.zeros((ne, 3, 3))                               # shape_funcs.py:369-374
    E[:, 0, 0] = sq12 * (sq12 - x3 + 3 * sq31 - y3) / (12 * J)
    E[:, 1, 1] = sq23 * (sq12 + xx + sq31 + yy) / (12 * J)
    E[:, 2, 2] = sq31 * (3 * sq12 - x3 + sq31 - y3) / (12 * J)
    E[:, 0, 1] = E[:, 1, 0] = l12 * l23 * (sq12 - xx - sq31 - yy) / (12 * J)
    E[:, 1, 2] = E[:, 2, 1] = l23 * l31 * (-sq12 - xx + sq31 - yy) / (12 * J)
    E[:, 2, 0] = E[:, 0, 2] = l31 * l12 * (-sq12 + x3 - sq31 + y3) / (12 * J)

    ls = np.stack([l12, l23, l31], axis=1)                 # shape_funcs.py:382-383
    C = (2 / J)[:, None, None] * ls[:, :, None] * ls[:, None, :]

    D = np.zeros((ne, 3, 3))                               # shape_funcs.py:393-401
    D[:, 0, 0] = l12 * (-sq12 + x3 - 2 * sq31 + y3) / (6 * J)
    D[:, 1, 1] = l23 * (-xx - sq31 - yy) / (6 * J)
    D[:, 2, 2] = l31 * (-2 * sq12 + xx + yy) / (6 * J)
    D[:, 0, 1] = l12 * (-xx + 2 * sq31 - yy) / (6 * J)
    D[:, 1, 0] = l23 * (-sq12 + sq31) / (6 * J)
    D[:, 1, 2] = l23 * (sq12 + xx + yy) / (6 * J)
    D[:, 2, 1] = l31 * (2 * xx + 2 * yy - sq31) / (6 * J)
    D[:, 2, 0] = l31 * (2 * sq12 - x3 + sq31 - y3) / (6 * J)
    D[:, 0, 2] = l12 * (sq12 - 2 * xx - 2 * yy) / (6 * J)

    N = np.empty((ne, 3, 3))                               # shape_funcs.py:410-411
    N[:] = (J / 24)[:, None, None]
    for i in range(3):
        N[:, i, i] = J / 12

    G = np.zeros((ne, 3, 3))                               # shape_funcs.py:421-426
    G[:, 0, 0] = (sq12 + sq31 - 2 * xx - 2 * yy) / (2 * J)
    G[:, 1, 1] = sq31 / (2 * J)
    G[:, 2, 2] = sq12 / (2 * J)
    G[:, 0, 1] = G[:, 1, 0] = (-sq31 + xx + yy) / (2 * J)
    G[:, 1, 2] = G[:, 2, 1] = (-xx - yy) / (2 * J)
    G[:, 2, 0] = G[:, 0, 2] = (-sq12 + xx + yy) / (2 * J)
    return E, C, D, N, G


def _sum_in_order(n_rows, n_cols, rows, cols, vals):
    """Sum duplicate (row, col) entries *sequentially in the order given* (the order of
    the reference's element loop, fe_solver.py:182-199) and drop results that are exactly
    0.0 (lil_matrix deletes zero entries on every write).  Returns a canonical CSC matrix."""
    key = cols.astype(np.int64) * n_rows + rows.astype(np.int64)      # CSC order
    order = np.argsort(key, kind="stable")
    ks = key[order]
    vs = vals[order]
    if len(ks) == 0:
        return sp.csc_matrix((n_rows, n_cols))
    start = np.r_[True, ks[1:] != ks[:-1]]
    seg = np.cumsum(start) - 1
    first = np.nonzero(start)[0]
    rank = np.arange(len(ks)) - first[seg]
    acc = np.zeros(len(first))
    for r in range(int(rank.max()) + 1):
        m = rank == r
        acc[seg[m]] = acc[seg[m]] + vs[m]
    uk = ks[first]
    nz = acc != 0.0
    uk, acc = uk[nz], acc[nz]
    r_, c_ = uk % n_rows, uk // n_rows
    indptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(indptr, c_ + 1, 1)
    indptr = np.cumsum(indptr)
    idt = np.int32 if max(n_rows, len(uk)) < 2 ** 31 else np.int64
    return sp.csc_matrix((acc, r_.astype(idt), indptr.astype(idt)), shape=(n_rows, n_cols))


def construct_AB_vec(mesh, IOR_dict, k):
    """Vector (Whitney edge + P1 node) assembly, sparse=True (fe_solver.py:148-214).
    DOFs: edges [0,Ntt) then nodes [Ntt,Ntt+Nzz) (fe_solver.py:159-162, 201-210).
    A = [[Att,0],[0,0]], B = [[Btt,Btz],[Btz^T,Bzz]] with
    Att += k^2n^2 NeNe - curlcurl; Btt += NeNe; Btz += NedN; Bzz += dNdN - k^2n^2 NN
    (fe_solver.py:196-199).  Returned as canonical CSC without stored zeros
    (fe_solver.py:212-213)."""
    conn, mat, labels = material_major(mesh)
    eidx, _, _ = material_major(mesh, np.asarray(mesh.edge_indices))
    conn = conn.astype(np.int64)
    eidx = eidx.astype(np.int64)
    Ntt = len(mesh.cells[0].data)
    Nzz = len(mesh.points)
    N = Ntt + Nzz
    E, C, D, NNl, G = vec_element_blocks(mesh.points[conn], conn)
    k2 = np.empty(len(conn))
    for mi, label in enumerate(labels):
        k2[mat == mi] = k ** 2 * IOR_dict[label] ** 2                 # fe_solver.py:181
    k2 = k2[:, None, None]
    Att = k2 * E - C
    Bzz = G - k2 * NNl
    # local (i,j) -> global indices, np.ix_ semantics (fe_solver.py:192-194)
    ri = np.repeat(eidx, 3, 1).ravel()
    cj = np.tile(eidx, (1, 3)).ravel()
    rz = np.repeat(conn, 3, 1).ravel() + Ntt
    cz = np.tile(conn, (1, 3)).ravel() + Ntt
    A = _sum_in_order(N, N, ri, cj, Att.ravel())
    rows = np.concatenate([ri, ri, cz, rz])
    cols = np.concatenate([cj, cz, ri, cz])
    vals = np.concatenate([E.ravel(), D.ravel(), D.ravel(), Bzz.ravel()])
    # element order must be preserved per slot; blocks never share a slot so the
    # concatenation keeps the per-slot order.
    B = _sum_in_order(N, N, rows, cols, vals)
    return A, B


def construct_AB_order1(mesh, IOR_dict, k):
    """Scalar P1 assembly, sparse=True (fe_solver.py:103-129): per element
    A[ix] += (k^2 n^2) LNN - LdNdN, B[ix] += LNN into N x N lil matrices, N = len(points)
    (fe_solver.py:107); element matrices from shape_funcs.py:403-428.  Returned as canonical
    CSC without stored zeros (what lil -> tocsr/tocsc holds)."""
    conn, mat, labels = material_major(mesh)
    conn = conn.astype(np.int64)
    N = len(mesh.points)
    _, _, _, NNl, G = vec_element_blocks(mesh.points[conn], conn)
    k2 = np.empty(len(conn))
    for mi, label in enumerate(labels):
        k2[mat == mi] = k ** 2 * IOR_dict[label] ** 2
    Ae = k2[:, None, None] * NNl - G
    r = np.repeat(conn, 3, 1).ravel()
    c = np.tile(conn, (1, 3)).ravel()
    return _sum_in_order(N, N, r, c, Ae.ravel()), _sum_in_order(N, N, r, c, NNl.ravel())


# --------------------------------------------------------------------------------------
# solves
# --------------------------------------------------------------------------------------
def count_modes_scalar(w_desc, k, IOR_dict):
    """fe_solver.py:330-346: walk eigenvalues descending, skip negatives, stop at the first
    neff outside [nmin, nmax]."""
    nmin, nmax = min(IOR_dict.values()), max(IOR_dict.values())
    cnt = 0
    for w in w_desc:
        if w < 0:
            continue
        ne = np.sqrt(w / k ** 2)
        if nmin <= ne <= nmax:
            cnt += 1
        else:
            break
    return cnt


def count_modes_vec(w, k, IOR_dict, target_neff):
    """fe_solver.py:409-424: same walk but in the order returned by the eigensolver, and
    every non-negative eigenvalue counts when target_neff is given."""
    nmin, nmax = min(IOR_dict.values()), max(IOR_dict.values())
    cnt = 0
    for _w in w:
        if _w < 0:
            continue
        ne = np.sqrt(_w / k ** 2)
        if (nmin <= ne <= nmax) or target_neff is not None:
            cnt += 1
        else:
            break
    return cnt


def solve_waveguide(mesh, wl, IOR_dict, Nmax=10, target_neff=None, AB=None, order=2):
    """Scalar solve, sparse=True (fe_solver.py:292-348): sigma = (k n_max)^2 or
    (k target)^2, eigsh(A, M=B, k=Nmax, which='SA', sigma=sigma), results reversed to
    descending order."""
    k = 2 * np.pi / wl
    sigma = np.power(k * (max(IOR_dict.values()) if target_neff is None else target_neff), 2)
    if AB is None:
        AB = construct_AB_order2(mesh, IOR_dict, k) if order == 2 else construct_AB_order1(mesh, IOR_dict, k)
    A, B = AB
    w, v = spla.eigsh(A.tocsr(), M=B.tocsr(), k=Nmax, which="SA", sigma=sigma)
    return w[::-1], v.T[::-1], count_modes_scalar(w[::-1], k, IOR_dict)


def solve_sparse(A, B, mesh, k, IOR_dict, num_modes=6):
    """fe_solver.py:262-290: the eigsh call of solve_waveguide on matrices the caller brings,
    sigma always (k n_max)^2, default 6 modes; results reversed to descending order."""
    sigma = np.power(k * max(IOR_dict.values()), 2)
    w, v = spla.eigsh(A.tocsr(), M=B.tocsr(), k=num_modes, which="SA", sigma=sigma)
    return w[::-1], v.T[::-1], count_modes_scalar(w[::-1], k, IOR_dict)


def compute_IOR_arr(mesh, IOR_dict):
    """fe_solver.py:446-452: index per element in cells[1].data order; unassigned elements stay 0,
    later material sets overwrite earlier ones."""
    out = np.zeros(len(mesh.cells[1].data))
    for material, sel in mesh.cell_sets.items():
        out[np.asarray(sel[1], dtype=np.int64)] = IOR_dict[material]
    return out


def solve_waveguide_vec(mesh, wl, IOR_dict, Nmax=10, target_neff=None, AB=None, return_info=False):
    """Vector solve, 'transform' semantics (fe_solver.py:397-400) applied matrix-free:
    the reference forms the dense C = (A - sigma B)^-1 B with spsolve and calls
    eigs(C, Nmax, which='SR'); here C is a LinearOperator x -> splu(A - sigma B).solve(B x),
    which has the same spectrum and needs O(nnz) memory.  w = sigma + 1/mu."""
    k = 2 * np.pi / wl
    sigma = np.power(k * (max(IOR_dict.values()) if target_neff is None else target_neff), 2)
    A, B = construct_AB_vec(mesh, IOR_dict, k) if AB is None else AB
    N = A.shape[0]
    lu = spla.splu((A - sigma * B).tocsc())
    Bc = B.tocsr()
    nops = [0]

    def mv(x):
        nops[0] += 1
        return lu.solve(Bc @ x)

    C = spla.LinearOperator((N, N), matvec=mv, dtype=np.float64)
    mu, v = spla.eigs(C, Nmax, which="SR")
    w = sigma + 1 / mu
    cnt = count_modes_vec(w, k, IOR_dict, target_neff)
    if return_info:
        return w, v.T, cnt, {"n_op": nops[0], "nnz_lu": lu.L.nnz + lu.U.nnz}
    return w, v.T, cnt


def get_eff_index(wl, w):
    """fe_solver.py:431-434."""
    k = 2 * np.pi / wl
    return np.sqrt(w / k ** 2)


# ---------------------------------------------------------------------------------------------
# mode resampling: point location, weights, grid interpolation
# ---------------------------------------------------------------------------------------------
def _isinside_all(px, py, P0, P1, P2):
    """fe_solver.py:597-612 (include_vertex=True) for one point against arrays of triangles; same
    operation order as the scalar Python code."""
    v1x, v1y = P1[:, 0] - P0[:, 0], P1[:, 1] - P0[:, 1]
    v2x, v2y = P2[:, 0] - P0[:, 0], P2[:, 1] - P0[:, 1]
    d12 = v1x * v2y - v1y * v2x
    with np.errstate(divide="ignore", invalid="ignore"):
        a = ((px * v2y - py * v2x) - (P0[:, 0] * v2y - P0[:, 1] * v2x)) / d12
        b = -((px * v1y - py * v1x) - (P0[:, 0] * v1y - P0[:, 1] * v1x)) / d12
        return (a >= 0) & (b >= 0) & (a + b <= 1)


def get_tri_idxs_points(mesh, px, py, maxr=0, kdtree_order=True):
    """Containing triangle of every point: the one with the nearest centroid (what the KD-tree search
    of fe_solver.py:848-863 finds first) or the first in mesh order (fe_solver.py:615-633); -1 if none
    or beyond maxr.  Brute force over all triangles, chunked."""
    pts = np.asarray(mesh.points, dtype=np.float64)
    tris = np.asarray(mesh.cells[1].data)[:, :3].astype(np.int64)
    P0, P1, P2 = pts[tris[:, 0], :2], pts[tris[:, 1], :2], pts[tris[:, 2], :2]
    cen = np.mean(pts[tris][:, :3, :2], axis=1)                    # mesher.py:239
    out = np.full(len(px), -1, dtype=int)
    for i, (x, y) in enumerate(zip(px, py)):
        if maxr > 0 and np.sqrt(x ** 2 + y ** 2) > maxr:           # fe_solver.py:849-851
            continue
        hit = np.nonzero(_isinside_all(x, y, P0, P1, P2))[0]
        if len(hit) == 0:
            continue
        if kdtree_order:
            d = (x - cen[hit, 0]) ** 2 + (y - cen[hit, 1]) ** 2
            out[i] = hit[np.argmin(d)]
        else:
            out[i] = hit[0]
    return out


def get_tri_idxs_KDtree(mesh, xa, ya, maxr=0):
    """fe_solver.py:865-872."""
    X, Y = np.meshgrid(np.asarray(xa, dtype=np.float64), np.asarray(ya, dtype=np.float64), indexing="ij")
    return get_tri_idxs_points(mesh, X.ravel(), Y.ravel(), maxr, True).reshape(len(xa), len(ya))


def get_tri_idxs(mesh, xa, ya):
    """fe_solver.py:662-668."""
    X, Y = np.meshgrid(np.asarray(xa, dtype=np.float64), np.asarray(ya, dtype=np.float64), indexing="ij")
    return get_tri_idxs_points(mesh, X.ravel(), Y.ravel(), 0, False).reshape(len(xa), len(ya))


def _affine_uv(mesh, X, Y, T):
    """shape_funcs.py:36-49 for arrays of points X, Y inside triangles T (valid entries only)."""
    pts = np.asarray(mesh.points, dtype=np.float64)
    tris = np.asarray(mesh.cells[1].data).astype(np.int64)
    t = tris[T]
    P0, P1, P2 = pts[t[:, 0], :2], pts[t[:, 1], :2], pts[t[:, 2], :2]
    x21, y21 = P1[:, 0] - P0[:, 0], P1[:, 1] - P0[:, 1]
    x31, y31 = P2[:, 0] - P0[:, 0], P2[:, 1] - P0[:, 1]
    J = x21 * y31 - x31 * y21
    m00, m01, m10, m11 = y31 / J, (-x31) / J, (-y21) / J, x21 / J
    dx, dy = X - P0[:, 0], Y - P0[:, 1]
    return m00 * dx + m01 * dy, m10 * dx + m11 * dy, (t, P0, P1, P2, x21, y21, x31, y31, J)


def get_interp_weights(mesh, xa, ya, tri_idxs, order=2):
    """fe_solver.py:670-696 with shape_funcs.py:5-10 / 286-288."""
    n = 6 if order == 2 else 3
    X, Y = np.meshgrid(np.asarray(xa, dtype=np.float64), np.asarray(ya, dtype=np.float64), indexing="ij")
    w = np.full((len(xa), len(ya), n), np.nan)
    ok = tri_idxs != -1
    u, v, _ = _affine_uv(mesh, X[ok], Y[ok], tri_idxs[ok])
    if order == 2:
        cols = [2 * (1 - u - v) * (0.5 - u - v), 2 * u * (u - 0.5), 2 * v * (v - 0.5),
                4 * u * (1 - u - v), 4 * u * v, 4 * v * (1 - u - v)]
    else:
        cols = [1 - u - v, u, v]
    w[ok] = np.stack(cols, axis=-1)
    return w


def get_interp_weights_vec(mesh, xa, ya, tri_idxs):
    """fe_solver.py:698-717 with shape_funcs.py:331-353 (signs: shape_funcs.py:317-322)."""
    X, Y = np.meshgrid(np.asarray(xa, dtype=np.float64), np.asarray(ya, dtype=np.float64), indexing="ij")
    w = np.full((len(xa), len(ya), 3, 2), np.nan)
    ok = tri_idxs != -1
    x, y = X[ok], Y[ok]
    _, _, (t, P0, P1, P2, x21, y21, x31, y31, J) = _affine_uv(mesh, x, y, tri_idxs[ok])
    x32, y32 = P2[:, 0] - P1[:, 0], P2[:, 1] - P1[:, 1]
    l12 = np.sqrt(x21 * x21 + y21 * y21) * np.where(t[:, 0] < t[:, 1], 1, -1)
    l23 = np.sqrt(x32 * x32 + y32 * y32) * np.where(t[:, 1] < t[:, 2], 1, -1)
    l31 = np.sqrt(x31 * x31 + y31 * y31) * np.where(t[:, 2] < t[:, 0], 1, -1)
    x1, y1 = P0[:, 0], P0[:, 1]
    vals = np.empty((len(x), 3, 2))
    vals[:, 0, 0] = (-y + y31 + y1) / J * l12
    vals[:, 0, 1] = (x - x31 - x1) / J * l12
    vals[:, 1, 0] = (-y + y1) / J * l23
    vals[:, 1, 1] = (x - x1) / J * l23
    vals[:, 2, 0] = (-y + y21 + y1) / J * l31
    vals[:, 2, 1] = (x - x21 - x1) / J * l31
    w[ok] = vals
    return w


def interpolate(v, mesh, xa, ya, tri_idxs=None, interp_weights=None, order=2, maxr=0):
    """fe_solver.py:719-744."""
    tri_idxs = get_tri_idxs_KDtree(mesh, xa, ya, maxr) if tri_idxs is None else tri_idxs
    interp_weights = get_interp_weights(mesh, xa, ya, tri_idxs, order) if interp_weights is None else interp_weights
    tris = np.asarray(mesh.cells[1].data).astype(np.int64)
    return np.sum(np.asarray(v)[tris[tri_idxs]] * interp_weights, axis=2)


def interpolate_vec(v, mesh, xa, ya, tri_idxs=None, interp_weights=None, maxr=0):
    """fe_solver.py:746-774."""
    Ntt = len(mesh.cells[0].data)
    vtt = np.asarray(v)[:Ntt]
    tri_idxs = get_tri_idxs_KDtree(mesh, xa, ya, maxr) if tri_idxs is None else tri_idxs
    interp_weights = get_interp_weights_vec(mesh, xa, ya, tri_idxs) if interp_weights is None else interp_weights
    field_points = vtt[np.asarray(mesh.edge_indices).astype(np.int64)[tri_idxs]]
    return np.sum(field_points[:, :, :, None] * interp_weights, axis=2)


def interpolate_kdtree_loop(v, mesh, xa, ya, maxr=0):
    """The reference's own control flow for a P2 mode (fe_solver.py:719-744 with the per-point KD-tree
    search of fe_solver.py:848-872 and the per-point weight loop of fe_solver.py:670-696), kept as a
    loop: this is what bench / tools time as the CPU baseline of the resampling path."""
    from scipy.spatial import KDTree
    pts = np.asarray(mesh.points, dtype=np.float64)
    tris = np.asarray(mesh.cells[1].data).astype(np.int64)
    tree = KDTree(np.mean(pts[tris][:, :3, :2], axis=1))                       # mesher.py:235-240
    out = np.full((len(xa), len(ya)), np.nan, dtype=np.asarray(v).dtype)
    max_tries = min(pts.shape[0] - 1, len(tris))
    for i in range(len(xa)):
        for j in range(len(ya)):
            p = np.array([xa[i], ya[j]])
            if maxr > 0 and np.sqrt(p[0] ** 2 + p[1] ** 2) > maxr:
                continue
            for tryno in range(max_tries):
                e = tree.query(p, [tryno + 1])[1][0]
                T = pts[tris[e, :3], :2]
                if _isinside_all(p[0], p[1], T[0:1], T[1:2], T[2:3])[0]:
                    x21, y21 = T[1, 0] - T[0, 0], T[1, 1] - T[0, 1]
                    x31, y31 = T[2, 0] - T[0, 0], T[2, 1] - T[0, 1]
                    J = x21 * y31 - x31 * y21
                    u, w = np.dot(np.array([[y31, -x31], [-y21, x21]]) / J, p - T[0])
                    N = np.array([2 * (1 - u - w) * (0.5 - u - w), 2 * u * (u - 0.5), 2 * w * (w - 0.5),
                                  4 * u * (1 - u - w), 4 * u * w, 4 * w * (1 - u - w)])
                    out[i, j] = np.sum(np.asarray(v)[tris[e]] * N)
                    break
    return out


# ---------------------------------------------------------------------------------------------
# overlaps in the mass-matrix inner product
# ---------------------------------------------------------------------------------------------
def overlap_matrix(U, V, B, conj=False):
    """`u.T @ B @ v` (notes eq. 38) for stacks of modes stored as rows -- the expression a reference
    user writes with B = construct_B(mesh, sparse=True) (fe_solver.py:71-101)."""
    U, V = np.asarray(U), np.asarray(V)
    return (U.conj() if conj else U) @ (B @ V.T)


def normalize_modes(U, B):
    """u / sqrt(u^H B u) per row: the normalisation eigsh applies to the vectors it returns (fe_solver.py:328)."""
    U = np.atleast_2d(np.asarray(U))
    d = np.einsum("in,ni->i", U.conj(), B @ U.T).real
    return U / np.sqrt(d)[:, None]
