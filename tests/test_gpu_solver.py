"""GPU: the sparse direct solver (K8) through the C ABI -- factor + solve against scipy's
SuperLU on the same matrices, pivot inertia, the quasi-definite form of the vector problem."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

K0 = 2 * np.pi / 1.55


@pytest.fixture(scope="module")
def ws():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import wavesolve_b200
    return wavesolve_b200


def _symbolic(sv, prob, fcs, leaf, where):
    if where == "host":
        return sv.Symbolic(prob.h_dofs, prob.h_xy, prob.N, fcs, leaf)
    return sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, fcs, leaf)


def _csr(prob, vals):
    indptr, indices = prob.pattern.export()
    return sp.csr_matrix((vals.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(prob.N,) * 2)


@pytest.mark.parametrize("where", ["device", "host"])
@pytest.mark.parametrize("ne,leaf,target", [(800, 16, None), (20000, 48, None), (60000, 48, None), (20000, 32, 1.447)])
def test_scalar_factor_solve(ws, ne, leaf, target, where):
    from wavesolve_b200 import fe_solver as fs, solver as sv
    mg = ws.meshgen
    mesh = mg.fiber_disc(ne, seed=3)
    prob = fs.ScalarProblem(mesh, mg.FIBER_IOR, K0)
    A, B = prob.assemble()
    sigma = (K0 * (max(mg.FIBER_IOR.values()) if target is None else target)) ** 2
    sym = _symbolic(sv, prob, 0, leaf, where)
    S = sv.Solver(prob.pattern, sym)
    st = S.factor(A, B, sigma)
    M = (_csr(prob, A) - sigma * _csr(prob, B)).tocsc()
    rng = np.random.default_rng(0)
    b = rng.standard_normal(prob.N)
    x = S.solve(torch.from_numpy(b).cuda()).cpu().numpy()
    res = np.linalg.norm(M @ x - b) / np.linalg.norm(b)
    xr = spla.splu(M).solve(b)
    assert res <= 1e-11, res
    assert np.linalg.norm(x - xr) <= 1e-9 * np.linalg.norm(xr)
    assert st["bad_pivots"] == 0 and st["neg_pivots"] + st["pos_pivots"] == prob.N
    if target is None:
        assert st["pos_pivots"] == 0          # sigma*B - A is SPD at the default shift
    else:
        # Sylvester: positive pivots = eigenvalues of (A,B) above sigma
        w = spla.eigsh(_csr(prob, A), M=_csr(prob, B), k=6, which="SA", sigma=(K0 * 1.46) ** 2)[0]
        assert st["pos_pivots"] == int((w > sigma).sum())
    # two right-hand sides in one call
    b2 = rng.standard_normal((2, prob.N))
    x2 = S.solve(torch.from_numpy(b2).cuda()).cpu().numpy()
    for i in range(2):
        assert np.linalg.norm(M @ x2[i] - b2[i]) <= 1e-11 * np.linalg.norm(b2[i])


@pytest.mark.parametrize("where", ["device", "host"])
@pytest.mark.parametrize("gen,arg,target", [("fiber_disc", 20000, None), ("structured_square", 60, None),
                                            ("fiber_disc", 20000, 1.448)])
def test_vector_quasidefinite_factor_solve(ws, gen, arg, target, where):
    from wavesolve_b200 import fe_solver as fs, solver as sv
    mg = ws.meshgen
    mesh = getattr(mg, gen)(arg, order=1)
    ws.get_unique_edges(mesh)
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, K0)
    A, B = prob.assemble()
    sigma = (K0 * (max(mg.FIBER_IOR.values()) if target is None else target)) ** 2
    Am, Bm = _csr(prob, A), _csr(prob, B)
    M = (Am - sigma * Bm).tocsc()
    Ntt, N = prob.Ntt, prob.N
    # the closed-form quasi-definite matrix equals T^T M T
    edges = prob.h_edges.astype(np.int64)
    L = np.linalg.norm(prob.h_xy[edges[:, 1]] - prob.h_xy[edges[:, 0]], axis=1)
    G = sp.csr_matrix((np.r_[1 / L, -1 / L], (np.r_[np.arange(Ntt), np.arange(Ntt)], np.r_[edges[:, 1], edges[:, 0]])),
                      shape=(Ntt, N - Ntt))
    T = sp.bmat([[sp.identity(Ntt), -G], [None, sp.identity(N - Ntt)]]).tocsr()
    Mq = prob.assemble_qd(sigma)
    Mq_ref = (T.T @ M @ T).tocsr()
    Mq_dev = _csr(prob, Mq)
    assert abs(Mq_dev - Mq_ref).max() <= 1e-11 * abs(Mq_ref).max()
    # T and T^T on the device
    v = np.random.default_rng(1).standard_normal(N)
    vd = torch.from_numpy(v).cuda()
    assert np.allclose(prob.apply_T(vd).cpu().numpy(), T @ v, rtol=1e-13, atol=1e-13)
    assert np.allclose(prob.apply_T(vd, transpose=True).cpu().numpy(), T.T @ v, rtol=1e-13, atol=1e-13)
    # factor M' (nodes first inside every front), solve M x = b as x = T M'^-1 T^T b
    sym = _symbolic(sv, prob, Ntt, 32, where)
    S = sv.Solver(prob.pattern, sym)
    st = S.factor(Mq)
    b = np.random.default_rng(2).standard_normal(N)
    y = S.solve(prob.apply_T(torch.from_numpy(b).cuda(), transpose=True))
    x = prob.apply_T(y).cpu().numpy()
    res = np.linalg.norm(M @ x - b) / np.linalg.norm(b)
    assert res <= 1e-10, res
    assert st["bad_pivots"] == 0
    if target is None:
        assert st["neg_pivots"] == Ntt and st["pos_pivots"] == N - Ntt


def test_solver_rejects_use_before_factor(ws):
    from wavesolve_b200 import fe_solver as fs, solver as sv, _lib
    mg = ws.meshgen
    prob = fs.ScalarProblem(mg.structured_square(6), mg.FIBER_IOR, K0)
    sym = sv.Symbolic(prob.h_dofs, prob.h_xy, prob.N, 0, 8)
    S = sv.Solver(prob.pattern, sym)
    with pytest.raises(_lib.WavesolveB200Error):
        S.solve(torch.zeros(prob.N, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("kind,ne,leaf", [("scalar", 3000, 16), ("vector", 20000, 32), ("vector", 40, 48), ("scalar", 150000, 48)])
def test_device_symbolic_structure(ws, kind, ne, leaf):
    """The device analysis (K8a on the GPU) yields a valid nested dissection: a permutation, the
    same balanced tree shape as the host model, sorted boundaries made of ancestor-owned
    unknowns, parent-relative positions that point at the same unknown, and a fill within a few
    percent of the host analysis of the same mesh (ties at the k-d splits are broken differently)."""
    from wavesolve_b200 import fe_solver as fs, solver as sv
    mg = ws.meshgen
    if kind == "scalar":
        prob = fs.ScalarProblem(mg.fiber_disc(ne, seed=5), mg.FIBER_IOR, K0)
        fcs = 0
    else:
        mesh = mg.fiber_disc(ne, order=1, seed=5)
        ws.get_unique_edges(mesh)
        prob = fs.VectorProblem(mesh, mg.FIBER_IOR, K0)
        fcs = prob.Ntt
    d = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, fcs, leaf)
    h = sv.Symbolic(prob.h_dofs, prob.h_xy, prob.N, fcs, leaf)
    assert (d.N, d.depth, d.nfronts) == (h.N, h.depth, h.nfronts)
    assert abs(d.nnzL - h.nnzL) <= 0.05 * h.nnzL + 1000
    E = d.export()
    perm, p, q, start, bndoff, bnd, rel = (E[k] for k in ("perm", "p", "q", "start", "bndoff", "bnd", "rel"))
    N = prob.N
    assert np.array_equal(np.sort(perm), np.arange(N))
    assert int(p[1:].sum()) == N and p[0] == 0
    # fronts own contiguous, level-major (leaves first) index ranges
    order = np.concatenate([np.arange(1 << l, 1 << (l + 1)) for l in range(d.depth, -1, -1)])
    assert np.array_equal(start[order], np.concatenate([[0], np.cumsum(p[order])[:-1]]))
    iperm = np.empty(N, dtype=np.int64)
    iperm[perm] = np.arange(N)
    owner = np.repeat(order, p[order])                      # owner front of every new index
    assert np.array_equal(np.diff(bndoff)[1:], q[1:]) and q[1] == 0
    # every element's unknowns are owned by ancestors-or-self of one leaf (its own)
    dofs = prob.h_dofs.astype(np.int64)
    own = owner[iperm[dofs]]                                # (Ne, 6) owner fronts
    lvl = np.floor(np.log2(own)).astype(np.int64)
    deepest = own[np.arange(len(own)), lvl.argmax(axis=1)]
    for a in range(6):
        sh = lvl.max(axis=1) - lvl[:, a]
        assert np.array_equal(deepest >> sh, own[:, a])
    for t in range(1, d.nfronts):
        b = bnd[bndoff[t]:bndoff[t + 1]]
        if len(b) == 0:
            continue
        assert np.all(np.diff(b) > 0)
        ob = owner[b]
        lt = int(np.log2(t))
        lo = np.floor(np.log2(ob)).astype(np.int64)
        assert np.all(lo < lt) and np.array_equal(t >> (lt - lo), ob)      # owned by proper ancestors
        par = t >> 1
        front_par = np.concatenate([np.arange(start[par], start[par] + p[par]), bnd[bndoff[par]:bndoff[par + 1]]])
        assert np.array_equal(front_par[rel[bndoff[t]:bndoff[t + 1]]], b)


def test_flag_synchronised_sweeps_are_bitwise_reproducible(ws):
    """The solve sweeps synchronise fronts through per-front completion flags and programmatic dependent launch
    (compute-sanitizer is closed on this pool, so this is the race check we can run): a missed dependency would read a
    stale remainder and change low-order bits from run to run.  200 solves of one right-hand side -- alone, with two
    right-hand sides per sweep, in plain stream order, and with four solvers interleaved on four streams -- must agree
    bit for bit, and the abort flag of the bounded spins must stay clear."""
    from wavesolve_b200 import fe_solver as fs, solver as sv, _lib, device as _dev
    mg = ws.meshgen
    mesh = mg.fiber_disc(60000, order=1, seed=14)
    ws.get_unique_edges(mesh)
    prob = fs.VectorProblem(mesh, mg.FIBER_IOR, K0)
    sigma = (K0 * max(mg.FIBER_IOR.values())) ** 2
    Mq = prob.assemble_qd(sigma)
    sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt)
    solvers = [sv.Solver(prob.pattern, sym) for _ in range(4)]
    for S in solvers:
        S.factor(Mq)
    b = torch.randn(prob.N, dtype=torch.float64, device="cuda")
    ref = solvers[0].solve(b).clone()
    M = _csr(prob, Mq)
    assert np.linalg.norm(M @ ref.cpu().numpy() - b.cpu().numpy()) <= 1e-10 * float(b.norm())
    x = torch.empty_like(b)
    for _ in range(200):
        solvers[0].solve(b, out=x)
        assert torch.equal(x, ref)
    b2 = torch.stack([b, b * 0.5])
    x2 = solvers[0].solve(b2)
    assert torch.equal(x2[0], ref)
    _lib.check(_lib.lib().ws_solver_set_overlap(solvers[1].handle, 0))       # plain stream order: same arithmetic
    assert torch.equal(solvers[1].solve(b), ref)
    streams = [torch.cuda.Stream() for _ in solvers]
    outs = [torch.empty_like(b) for _ in solvers]
    torch.cuda.synchronize()
    for rep in range(50):
        for S, st, o in zip(solvers, streams, outs):
            with torch.cuda.stream(st):
                S.solve(b, out=o)
    torch.cuda.synchronize()
    for S, o in zip(solvers, outs):
        assert torch.equal(o, ref)
        S.check()                                                             # no bounded spin gave up
        S.close()


_RETILE_CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import wavesolve_b200 as ws
from wavesolve_b200 import fe_solver as fs, solver as sv
mg = ws.meshgen
K0 = 2 * np.pi / 1.55
mesh = mg.fiber_disc(50000, order=1, seed=21)
ws.get_unique_edges(mesh)
prob = fs.VectorProblem(mesh, mg.FIBER_IOR, K0)
sigma = (K0 * max(mg.FIBER_IOR.values())) ** 2
Mq = prob.assemble_qd(sigma)
sym = sv.Symbolic.on_device(prob.dofs, prob.xy, prob.N, prob.Ntt)
S = sv.Solver(prob.pattern, sym)
S.factor(Mq)
g = torch.Generator(device="cuda").manual_seed(5)
b = torch.randn((2, prob.N), dtype=torch.float64, device="cuda", generator=g)
x1 = S.solve(b[0].contiguous())
x2 = S.solve(b)
np.save(sys.argv[2], np.stack([x1.cpu().numpy(), x2[0].cpu().numpy(), x2[1].cpu().numpy(), b[0].cpu().numpy(), b[1].cpu().numpy()]))
"""


def test_top_levels_through_tile_kernels_agree_with_front_kernels(ws, tmp_path):
    """Levels of few, wide fronts near the root of a small problem take the 256-thread tile kernels (LevelPlan::narrow_tiles)
    instead of one thread per front row; WS_TILE_TOP=0 keeps them in the one-CTA-per-front kernels.  The choice is made once
    per process, so each variant solves the same systems (one and two right-hand sides) in a process of its own: the
    solutions agree to rounding (different summation order, same factors) and both satisfy M x = b."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tag, top in (("tiles", "32"), ("fronts", "0")):
        path = str(tmp_path / ("x_%s.npy" % tag))
        env = dict(os.environ, WS_TILE_TOP=top)
        r = subprocess.run([sys.executable, "-c", _RETILE_CHILD, root, path], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(path))
    a, f = outs
    assert np.array_equal(a[3:], f[3:])                                   # same right-hand sides
    for i in range(3):
        assert np.linalg.norm(a[i] - f[i]) <= 1e-11 * np.linalg.norm(f[i]), i
    assert np.array_equal(a[0], a[1]) and np.array_equal(f[0], f[1])       # NR = 2 carries the same arithmetic per vector
