"""TEST INFRASTRUCTURE ONLY -- numpy model of the device sparse solver (csrc/ws_solver*.cu).

A small, slow, readable multifrontal LDL^T with the same structure as the CUDA solver:
element-based geometric nested dissection (k-d bisection of element centroids, complete
binary tree), unknown -> owner node = lowest common ancestor of the leaves of its elements,
level-major elimination order, dense fronts [own | boundary], no pivoting, and "selective
inversion" (the solve sweeps use L11^-1 and L21 L11^-1 so that they are matrix-vector
products).  Used by the CPU tests to validate the host symbolic phase of the library and as
the specification the CUDA kernels are checked against.  Not a reference restatement: the
reference delegates this step to SuperLU (scipy splu/spsolve; fe_solver.py:328, :398).
"""
from __future__ import annotations

import numpy as np


def kd_leaves(cxy, depth):
    """leaf index of every element: recursive median bisection along the longer bbox axis."""
    ne = len(cxy)
    order = np.arange(ne)
    leaf = np.zeros(ne, dtype=np.int64)

    def rec(lo, hi, level, idx):
        if level == depth:
            leaf[order[lo:hi]] = idx
            return
        seg = order[lo:hi]
        if len(seg):
            p = cxy[seg]
            ax = 0 if (np.ptp(p[:, 0]) >= np.ptp(p[:, 1])) else 1
            mid = len(seg) // 2
            part = np.argpartition(p[:, ax], mid) if 0 < mid < len(seg) else np.arange(len(seg))
            order[lo:hi] = seg[part]
        mid = (lo + hi) // 2
        rec(lo, mid, level + 1, 2 * idx)
        rec(mid, hi, level + 1, 2 * idx + 1)

    rec(0, ne, 0, 0)
    return leaf


class Structure:
    pass


def analyse(dofs, cxy, N, leaf_elems=32):
    dofs = np.asarray(dofs, dtype=np.int64)
    ne = len(dofs)
    depth = 0
    while (ne >> depth) > leaf_elems:
        depth += 1
    leaf = kd_leaves(np.asarray(cxy), depth)
    lo = np.full(N, 1 << 40, dtype=np.int64)
    hi = np.full(N, -1, dtype=np.int64)
    for a in range(dofs.shape[1]):
        np.minimum.at(lo, dofs[:, a], leaf)
        np.maximum.at(hi, dofs[:, a], leaf)
    if (hi < 0).any():
        raise ValueError("unknowns without elements: matrix has empty rows")
    x = lo ^ hi
    hb = np.zeros(N, dtype=np.int64)          # number of differing low bits
    nz = x > 0
    hb[nz] = np.floor(np.log2(x[nz])).astype(np.int64) + 1
    node = ((1 << depth) + lo) >> hb           # heap id of the owner
    level = depth - hb
    # level-major bottom-up order, inside a node by original index
    key = (depth - level) * (1 << (depth + 2)) + node
    perm = np.lexsort((np.arange(N), key))     # new -> old
    iperm = np.empty(N, dtype=np.int64)
    iperm[perm] = np.arange(N)
    nnodes = 1 << (depth + 1)
    cnt = np.bincount(node, minlength=nnodes)
    start = np.zeros(nnodes, dtype=np.int64)
    pos = 0
    for l in range(depth, -1, -1):
        for t in range(1 << l, 1 << (l + 1)):
            start[t] = pos
            pos += cnt[t]
    # boundary lists
    bnd = [None] * nnodes
    new_dofs = iperm[dofs]
    for t in range(1 << depth, 1 << (depth + 1)):
        sel = leaf == (t - (1 << depth))
        u = np.unique(new_dofs[sel])
        bnd[t] = u[u >= start[t] + cnt[t]]
        assert np.all((u >= start[t]))
    for l in range(depth - 1, -1, -1):
        for t in range(1 << l, 1 << (l + 1)):
            u = np.union1d(bnd[2 * t], bnd[2 * t + 1])
            own = (u >= start[t]) & (u < start[t] + cnt[t])
            assert own.sum() == cnt[t], "every own unknown must appear in a child boundary"
            rest = u[~own]
            assert np.all(rest >= start[t] + cnt[t])
            bnd[t] = rest
    S = Structure()
    S.N, S.depth, S.nnodes = N, depth, nnodes
    S.perm, S.iperm, S.start, S.cnt, S.bnd = perm, iperm, start, cnt, bnd
    S.node_of_new = node[perm]
    S.nnzL = int(sum(cnt[t] * (cnt[t] + len(bnd[t])) for t in range(1, nnodes)))
    S.flops = float(sum(cnt[t] ** 3 / 3 + cnt[t] ** 2 * len(bnd[t]) + cnt[t] * len(bnd[t]) ** 2
                        for t in range(1, nnodes)))
    return S


def structure_from_arrays(N, depth, perm, p, start, bndoff, bnd):
    """Wrap the arrays exported by the library's host symbolic phase (ws_symbolic_export) so
    that factor()/solve() of this model run on exactly that elimination tree."""
    S = Structure()
    S.N, S.depth, S.nnodes = int(N), int(depth), 1 << (int(depth) + 1)
    S.perm = np.asarray(perm, dtype=np.int64)
    S.iperm = np.empty(S.N, dtype=np.int64)
    S.iperm[S.perm] = np.arange(S.N)
    S.cnt = np.asarray(p, dtype=np.int64)
    S.start = np.asarray(start, dtype=np.int64)
    S.bnd = [None] + [np.asarray(bnd[bndoff[t]:bndoff[t + 1]], dtype=np.int64) for t in range(1, S.nnodes)]
    S.nnzL = int(sum(S.cnt[t] * (S.cnt[t] + len(S.bnd[t])) for t in range(1, S.nnodes)))
    return S


def ldl_nopiv(F, p):
    """in place: first p columns of the symmetric front F (m x m) eliminated without pivoting.
    Returns d (p,), L panel (m x p, unit diagonal implied), Schur complement (m-p x m-p)."""
    F = F.copy()
    m = F.shape[0]
    d = np.zeros(p)
    for c in range(p):
        d[c] = F[c, c]
        l = F[c + 1:, c] / d[c]
        F[c + 1:, c + 1:] -= np.outer(l, F[c + 1:, c])
        F[c + 1:, c] = l
    L = np.tril(F[:, :p], -1)
    L[np.arange(p), np.arange(p)] = 1.0
    return d, L, F[p:, p:]


def factor(S, M, invert=True):
    """M: scipy sparse symmetric (full pattern).  Returns per-node solve panels."""
    Mp = M.tocsr()[S.perm][:, S.perm].tocsc()
    U = [None] * S.nnodes
    panels = [None] * S.nnodes
    dall = np.zeros(S.N)
    for l in range(S.depth, -1, -1):
        for t in range(1 << l, 1 << (l + 1)):
            p, b = int(S.cnt[t]), S.bnd[t]
            idx = np.concatenate([np.arange(S.start[t], S.start[t] + p), b]).astype(np.int64)
            m = len(idx)
            F = np.zeros((m, m))
            if p:
                cols = Mp[:, S.start[t]:S.start[t] + p]
                sub = cols[idx].toarray()                  # m x p : original entries live in own columns
                F[:, :p] = sub
                F[:p, :] = sub.T
            if l < S.depth:
                for c in (2 * t, 2 * t + 1):
                    rel = np.searchsorted(idx, S.bnd[c])
                    assert np.array_equal(idx[rel], S.bnd[c])
                    F[np.ix_(rel, rel)] += U[c]
                    U[c] = None
            d, L, Sc = ldl_nopiv(F, p)
            U[t] = Sc
            dall[S.start[t]:S.start[t] + p] = d
            if invert and p:
                Linv = np.linalg.solve(L[:p], np.eye(p))
                W = L[p:] @ Linv
                panels[t] = np.vstack([np.tril(Linv), W])
            else:
                panels[t] = L
    S.panels, S.d, S.inverted = panels, dall, invert
    return S


def solve(S, b):
    assert S.inverted
    w = np.asarray(b, dtype=float)[S.perm].copy()
    y = np.zeros(S.N)
    upd = [None] * S.nnodes
    for l in range(S.depth, -1, -1):
        for t in range(1 << l, 1 << (l + 1)):
            p, bd = int(S.cnt[t]), S.bnd[t]
            idx = np.concatenate([np.arange(S.start[t], S.start[t] + p), bd]).astype(np.int64)
            wv = np.zeros(len(idx))
            wv[:p] = w[S.start[t]:S.start[t] + p]
            if l < S.depth:
                for c in (2 * t, 2 * t + 1):
                    rel = np.searchsorted(idx, S.bnd[c])
                    wv[rel] += upd[c]
                    upd[c] = None
            out = S.panels[t] @ wv[:p] if p else np.zeros(len(idx))
            y[S.start[t]:S.start[t] + p] = out[:p]
            upd[t] = wv[p:] - out[p:]
    x = np.zeros(S.N)
    for l in range(0, S.depth + 1):
        for t in range(1 << l, 1 << (l + 1)):
            p, bd = int(S.cnt[t]), S.bnd[t]
            if not p:
                continue
            g = np.concatenate([y[S.start[t]:S.start[t] + p] / S.d[S.start[t]:S.start[t] + p], -x[bd]])
            x[S.start[t]:S.start[t] + p] = S.panels[t].T @ g
    out = np.empty(S.N)
    out[S.perm] = x
    return out
