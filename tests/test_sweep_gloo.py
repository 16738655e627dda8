"""Host-side logic of the sweep path: partitioning and the eigenpair gather,
run on CPU with world_size 2 over gloo (the same code path NCCL takes on the GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wavesolve_b200 import sweep


def test_partition_covers_everything_once_and_is_deterministic():
    rng = np.random.default_rng(0)
    mesh_ids = list(rng.integers(0, 5, size=37))
    costs = [float(1 + (m + 1) ** 1.5) for m in mesh_ids]
    for world in (1, 2, 4, 8):
        t1 = sweep.partition(mesh_ids, costs, world)
        t2 = sweep.partition(mesh_ids, costs, world)
        assert t1 == t2
        assert sorted(i for r in t1 for i in r) == list(range(len(mesh_ids)))
        loads = [sum(costs[i] for i in r) for r in t1]
        assert max(loads) <= 1.5 * sum(costs) / world + max(costs)
        for r in t1:   # a mesh's problems are adjacent on a rank (context reuse)
            seen = []
            for i in r:
                if not seen or seen[-1] != mesh_ids[i]:
                    assert mesh_ids[i] not in seen
                    seen.append(mesh_ids[i])


def test_partition_splits_a_single_mesh_across_ranks():
    t = sweep.partition([0] * 16, [1.0] * 16, 8)
    assert [len(r) for r in t] == [2] * 8
    t = sweep.partition([0] * 3, [1.0] * 3, 8)
    assert sorted(len(r) for r in t) == [0] * 5 + [1] * 3


def test_problem_grid_is_mesh_major_and_carries_tags():
    from wavesolve_b200 import sweep
    ior = {"core": 1.45, "cladding": 1.44}
    wls = np.linspace(1.0, 1.8, 4)
    g = sweep.problem_grid(3, wls, ior, tags=[9.0, 12.5, 16.0])
    assert [p.mesh_id for p in g] == [0] * 4 + [1] * 4 + [2] * 4
    assert [p.wl for p in g[:4]] == [float(w) for w in wls] and g[5].tag == (12.5, float(wls[1]))
    assert all(p.IOR_dict == ior and p.IOR_dict is not ior for p in g)
    per_mesh = [{"core": 1.45 + 0.01 * m, "cladding": 1.44} for m in range(3)]
    g = sweep.problem_grid(3, [1.55], per_mesh, target_neff=1.444)
    assert [p.IOR_dict["core"] for p in g] == [1.45, 1.46, 1.47] and g[0].target_neff == 1.444 and g[2].tag == (2, 1.55)
    g = sweep.problem_grid(2, [1.0, 2.0], lambda m, wl: {"core": 1.45 - 0.01 * wl, "cladding": 1.44})
    assert [round(p.IOR_dict["core"], 6) for p in g] == [1.44, 1.43, 1.44, 1.43]
    assert sweep.problem_grid(0, wls, ior) == [] and sweep.problem_grid(2, [], ior) == []
    with pytest.raises(ValueError):
        sweep.problem_grid(2, wls, per_mesh)
    with pytest.raises(ValueError):
        sweep.problem_grid(2, wls, ior, tags=[1.0])
    # the grid feeds partition: the wavelengths of one mesh stay together on a rank
    table = sweep.partition([p.mesh_id for p in sweep.problem_grid(4, wls, ior)], [1.0] * 16, 2)
    assert sorted(len(t) for t in table) == [8, 8]


def test_default_lanes_by_problem_size():
    assert sweep.default_lanes(100_000) == 8 and sweep.default_lanes(500_000) == 2 and sweep.default_lanes(2_000_000) == 1


def test_partition_edge_cases():
    assert sweep.partition([], [], 4) == [[], [], [], []]
    assert sweep.partition([3], [2.0], 1) == [[0]]


def test_gather_single_rank_identity():
    loc = {2: (torch.arange(3, dtype=torch.float64), torch.ones(3, 5, dtype=torch.float64), 2),
           0: (torch.zeros(3, dtype=torch.float64), torch.full((3, 4), 7.0, dtype=torch.float64), 1),
           1: (torch.ones(3, dtype=torch.float64), torch.zeros(3, 6, dtype=torch.float64), 0)}
    out = sweep.gather_eigenpairs(loc, 3)
    assert [o[2] for o in out] == [1, 0, 2]
    assert out[0][1].shape == (3, 4) and np.all(out[0][1] == 7.0)
    assert out[2][1].shape == (3, 5) and np.all(out[2][0] == np.arange(3))
    out = sweep.gather_eigenpairs(loc, 3, with_vectors=False)
    assert out[1][1] is None and np.all(out[1][0] == 1.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_result(i, n, k=4):
    w = torch.arange(k, dtype=torch.float64) + 100.0 * i
    V = (torch.arange(k * n, dtype=torch.float64).reshape(k, n) + i) / (1.0 + i)
    return w, V, i % 3


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh_ids = [0, 0, 0, 1, 1, 2, 2, 2, 2]
        Ns = {0: 11, 1: 5, 2: 8}
        costs = [sweep.estimate_cost(Ns[m]) for m in mesh_ids]
        table = sweep.partition(mesh_ids, costs, world)
        # packed send buffer exactly like solve_sweep builds it (ragged N across meshes)
        k = 4
        total = max(sum(k * (1 + Ns[mesh_ids[i]]) for i in r) for r in table)
        buf = torch.zeros(total, dtype=torch.float64)
        local, off = {}, 0
        for i in table[rank]:
            n = Ns[mesh_ids[i]]
            w, V, cnt = _fake_result(i, n, k)
            wv = buf[off:off + k]
            Vv = buf[off + k:off + k + k * n].view(k, n)
            wv.copy_(w)
            Vv.copy_(V)
            off += k * (1 + n)
            local[i] = (wv, Vv, cnt, 3 * (i % 2))          # per-problem status travels with the results
        got = sweep.gather_eigenpairs(local, len(mesh_ids), packed=buf)
        ok = True
        for i, (w, v, cnt, owner, rc) in enumerate(got):
            we, Ve, ce = _fake_result(i, Ns[mesh_ids[i]], k)
            ok = ok and np.array_equal(w, we.numpy()) and np.array_equal(v, Ve.numpy()) and cnt == ce
            ok = ok and i in table[owner] and rc == 3 * (i % 2)
        # staging path (no packed buffer) and dst-only delivery (point-to-point to rank 0, exact sizes)
        st = {}
        got2 = sweep.gather_eigenpairs(local, len(mesh_ids), dst=0, stats=st)
        if rank == 0:
            ok = ok and all(np.array_equal(a[1], b[1]) and a[4] == b[4] for a, b in zip(got, got2))
            ok = ok and st["bytes"] == 8 * sum(k * (1 + Ns[m]) for m in mesh_ids)
        else:
            ok = ok and got2 is None and st["bytes"] == 0
        got3 = sweep.gather_eigenpairs(local, len(mesh_ids), dst=1, with_vectors=False)
        ok = ok and ((got3 is None) if rank == 0 else all(g[1] is None and np.array_equal(g[0], a[0]) for g, a in zip(got3, got)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.gpu
def test_sweep_matches_individual_solves():
    import wavesolve_b200 as ws
    mg = ws.meshgen
    meshes = [mg.fiber_disc(4000, order=1, seed=3), mg.fiber_disc(6000, order=1, seed=4)]
    for m in meshes:
        ws.get_unique_edges(m)
    wls = [1.45, 1.55, 1.65]
    probs = [sweep.SweepProblem(mi, wl, mg.FIBER_IOR) for mi in range(2) for wl in wls]
    res = sweep.solve_sweep(meshes, probs, Nmax=6, kind="vector")
    assert [r.index for r in res] == list(range(len(probs)))
    # concurrent lanes (own stream + solver storage + host thread each) must give the same answers
    res3 = sweep.solve_sweep(meshes, probs, Nmax=6, kind="vector", lanes=3)
    for a, b in zip(res, res3):
        assert a.index == b.index and a.mode_count == b.mode_count
        assert np.max(np.abs(a.w - b.w) / np.abs(a.w)) < 1e-10
    for r, pr in zip(res, probs):
        w, v, cnt = ws.solve_waveguide_vec(meshes[pr.mesh_id], pr.wl, pr.IOR_dict, Nmax=6, verbose=False)
        k = 2 * np.pi / pr.wl
        ne_s, ne_i = np.sqrt(r.w) / k, np.sqrt(w.real) / k
        assert np.max(np.abs(ne_s - ne_i) / ne_i) < 1e-10
        assert r.mode_count == cnt
        assert r.v.shape == v.shape
        assert np.allclose(np.linalg.norm(r.v, axis=1), 1.0, atol=1e-12)
    # and against the ORACLE directly (not only against this package's single solves): first and last problem
    from oracle import restated as R
    from tests import util
    for i in (0, len(probs) - 1):
        pr = probs[i]
        wr, vr, cr = R.solve_waveguide_vec(meshes[pr.mesh_id], pr.wl, pr.IOR_dict, Nmax=6)
        order = np.argsort(-wr.real, kind="stable")
        k = 2 * np.pi / pr.wl
        assert np.allclose(np.sqrt(res3[i].w) / k, np.sqrt(wr.real[order]) / k, rtol=1e-10, atol=0)
        assert res3[i].mode_count == R.count_modes_vec(wr.real[order], k, pr.IOR_dict, None)
        for grp in util.group_degenerate(wr.real[order], rtol=1e-5):
            if grp[-1] == 5:
                continue
            assert util.subspace_overlap(res3[i].v[grp], vr[order][grp].real) >= 1 - 1e-8


@pytest.mark.gpu
def test_sweep_scalar_reuses_context():
    import wavesolve_b200 as ws
    mg = ws.meshgen
    mesh = mg.fiber_disc(3000, order=2, seed=5)
    probs = [sweep.SweepProblem(0, wl, mg.FIBER_IOR) for wl in (1.3, 1.55)]
    res = sweep.solve_sweep([mesh], probs, Nmax=5, kind="scalar")
    for r, pr in zip(res, probs):
        w, v, cnt = ws.solve_waveguide(mesh, pr.wl, pr.IOR_dict, Nmax=5)
        assert np.max(np.abs(r.w - w) / np.abs(w)) < 1e-11
        assert r.mode_count == cnt


@pytest.mark.gpu
def test_sweep_reports_a_failing_problem_and_finishes_the_rest():
    """error code + partial results per problem.  One poisoned problem (NaN wavelength -> NaN
    matrix) fails on its own; its neighbours -- same mesh, same lanes -- are solved."""
    import wavesolve_b200 as ws
    mg = ws.meshgen
    mesh = mg.fiber_disc(4000, order=1, seed=3)
    ws.get_unique_edges(mesh)
    probs = [sweep.SweepProblem(0, wl, mg.FIBER_IOR) for wl in (1.5, float("nan"), 1.6, 1.55)]
    for lanes in (1, 4):
        res = sweep.solve_sweep([mesh], probs, Nmax=4, kind="vector", lanes=lanes)
        assert [r.rc == 0 for r in res] == [True, False, True, True]
        assert np.all(np.isnan(res[1].w)) and res[1].mode_count == 0 and "error" in res[1].info
        for r, pr in zip(res, probs):
            if r.rc == 0:
                w, v, cnt = ws.solve_waveguide_vec(mesh, pr.wl, pr.IOR_dict, Nmax=4, verbose=False)
                assert np.max(np.abs(np.sqrt(r.w) - np.sqrt(w.real)) / np.sqrt(w.real)) < 1e-10 and r.mode_count == cnt
