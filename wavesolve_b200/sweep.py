"""Parameter sweeps (wavelengths x z-slices x index sets): one eigenproblem per GPU at a time,
NCCL only to gather the eigenpairs at the end.

The reference runs sweeps as a serial notebook loop that re-meshes per z-slice, re-assembles
per wavelength and pickles per radius (lantern-demo.ipynb cell 4).  Here a sweep is a list of
independent ``SweepProblem``s:

* ``problem_grid``   meshes x wavelengths (x index sets) -> the flat problem list;
* ``partition``      cost-sorted, mesh-grouped longest-first assignment of problems to ranks
                     (pure host logic; identical on every rank, no communication);
* ``solve_sweep``    each rank keeps one resident context per mesh (pattern, incidence lists,
                     symbolic analysis, solver task plan and front storage) and for every
                     wavelength only refreshes k^2 n^2, re-assembles the values, refactors and
                     iterates (K11); the Ritz vectors are written by the last kernel of the
                     eigensolver directly into the rank's send buffer;
* ``gather_eigenpairs``  one collective over ``torch.distributed`` (NCCL on GPUs; the same code
                     runs on gloo/CPU tensors in the tests): metadata by all_gather_object,
                     payload by one padded all_gather.

A single mesh is never sharded across GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


@dataclass
class SweepProblem:
    mesh_id: int                 # index into the list of meshes handed to solve_sweep
    wl: float
    IOR_dict: Dict[str, float]
    target_neff: Optional[float] = None
    tag: object = None           # free-form label carried to the result (e.g. (radius, wl))


@dataclass
class SweepResult:
    index: int                   # position in the problem list
    w: np.ndarray                # (Nmax,) eigenvalues, descending
    v: Optional[np.ndarray]      # (Nmax, N) eigenvectors as rows (None when gathered without vectors)
    mode_count: int
    rank: int
    info: dict = field(default_factory=dict)
    rc: int = 0                  # 0 ok; 3: fewer than Nmax pairs converged (w, v hold what the iteration had);
    #                              other codes: the problem failed (w, v are NaN), ``info["error"]`` says why --
    #                              one failing problem never aborts the sweep


def problem_grid(n_meshes: int, wavelengths: Sequence[float], IOR_dicts, target_neff: Optional[float] = None,
                 tags: Optional[Sequence] = None) -> List[SweepProblem]:
    """The problem list of the reference's sweep loop (lantern-demo.ipynb cell 4: for each z-slice
    mesh, for each wavelength, solve) as a flat list, mesh-major so that the problems of one mesh are
    adjacent.  ``IOR_dicts``: one dict for all problems, one per mesh (sequence of length ``n_meshes``),
    or a callable ``(mesh_id, wl) -> dict`` for dispersive materials.  ``tags[m]`` labels mesh ``m``
    (e.g. the taper radius); each problem carries ``tag = (tags[m], wl)``."""
    if tags is not None and len(tags) != n_meshes:
        raise ValueError("tags: one per mesh")
    if not callable(IOR_dicts) and not isinstance(IOR_dicts, dict) and len(IOR_dicts) != n_meshes:
        raise ValueError("IOR_dicts: one dict, one dict per mesh, or a callable")
    out = []
    for m in range(n_meshes):
        for wl in wavelengths:
            wl = float(wl)
            if callable(IOR_dicts):
                ior = IOR_dicts(m, wl)
            elif isinstance(IOR_dicts, dict):
                ior = IOR_dicts
            else:
                ior = IOR_dicts[m]
            out.append(SweepProblem(m, wl, dict(ior), target_neff, (tags[m] if tags is not None else m, wl)))
    return out


def estimate_cost(n_unknowns: int) -> float:
    """Relative cost of one eigensolve: 2-D nested dissection factorisation ~ N^1.5."""
    return float(n_unknowns) ** 1.5


def partition(mesh_ids: Sequence[int], costs: Sequence[float], world: int) -> List[List[int]]:
    """Assign problem indices to ``world`` ranks.

    Problems of one mesh form a group (they share pattern / symbolic analysis / plan).  Groups
    are placed longest-first on the least-loaded rank; a group larger than the ideal share is
    split into contiguous chunks first so that a sweep over few meshes still uses every GPU.
    Deterministic (ties broken by index), so every rank computes the same table."""
    assert world >= 1 and len(mesh_ids) == len(costs)
    n = len(mesh_ids)
    total = float(sum(costs))
    ideal = total / world if world else total
    groups: Dict[int, List[int]] = {}
    for i, m in enumerate(mesh_ids):
        groups.setdefault(int(m), []).append(i)
    chunks: List[List[int]] = []
    for m in sorted(groups):
        idx = groups[m]
        gcost = sum(costs[i] for i in idx)
        nsplit = 1
        if ideal > 0 and gcost > ideal and len(idx) > 1:
            nsplit = min(len(idx), int(np.ceil(gcost / ideal)))
        # contiguous, near-equal chunks
        bounds = [round(j * len(idx) / nsplit) for j in range(nsplit + 1)]
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b > a:
                chunks.append(idx[a:b])
    chunks.sort(key=lambda c: (-sum(costs[i] for i in c), c[0]))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for c in chunks:
        r = min(range(world), key=lambda j: (load[j], j))
        out[r].extend(c)
        load[r] += sum(costs[i] for i in c)
    for r in range(world):
        out[r].sort(key=lambda i: (mesh_ids[i], i))      # keep a mesh's wavelengths adjacent
    assert sorted(i for r in out for i in r) == list(range(n))
    return out


# ---------------------------------------------------------------------------------------------
# gather (K12)
# ---------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist


def _to_host(t: torch.Tensor) -> np.ndarray:
    """Device -> host for the gather payload: page-locked block for moderate sizes, double-buffered staging into
    pageable memory for GBs (device.d2h_staged)."""
    if t.is_cuda and t.numel() * t.element_size() >= (1 << 20):
        from . import device as _dev
        return _dev.d2h_staged(t)
    return t.cpu().numpy()


def gather_eigenpairs(local: Dict[int, tuple], n_problems: int, group=None, with_vectors=True,
                      dst: Optional[int] = None, packed: Optional[torch.Tensor] = None, stats: Optional[dict] = None,
                      own_host: Optional[np.ndarray] = None):
    """Collect per-problem results from all ranks.

    ``local``: {problem index: (w tensor (k,), V tensor (k, N), mode_count int[, rc int])} -- tensors on the
    rank's device (CUDA for NCCL, CPU for gloo).  Returns a list of ``n_problems`` entries
    ``(w ndarray, v ndarray | None, mode_count, owner_rank, rc)`` on every rank (``dst=None``) or on rank ``dst``
    only (other ranks get None).  Without an initialised process group this is the single-rank identity.

    ``dst=None``: one padded ``all_gather`` (every rank needs everything).  ``dst=r``: every other rank sends its
    payload -- exactly its own size, no padding -- to ``r`` with point-to-point ``send`` / ``recv`` (NCCL over
    NVLink on GPUs), so only ``r`` receives and only ``r`` pays the device -> host copy (through pinned memory).
    ``packed``: the flat buffer [w_i | V_i ...] (in ``local``'s insertion order) the tensors of ``local`` are views
    of; it is sent as is (no staging copy).  ``stats``: if a dict, filled with ``bytes`` received / held on this
    rank and ``seconds`` of the collective + host copy.  ``own_host``: a host copy of this rank's packed payload, if the
    caller already has one; used instead of copying the rank's own block again.  (Streaming the own results to the host
    while later meshes are solved -- a helper thread, or a landing buffer page-locked in the background -- was measured
    and is no faster than the staged copy at the end: 48.9 / 51.0 against 53.0 problems/s; both contend with the
    launching thread for the driver.)"""
    import time as _time
    dist = _dist()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    keys = list(local)            # insertion order = packing order
    meta = [(i, int(local[i][0].numel()), int(local[i][1].shape[-1]) if with_vectors else 0, int(local[i][2]),
             int(local[i][3]) if len(local[i]) > 3 else 0) for i in keys]
    t_start = _time.perf_counter()
    if multi:
        metas = [None] * world
        dist.all_gather_object(metas, meta, group=group)
    else:
        metas = [meta]
    sizes = [sum(k * (1 + n) for (_, k, n, _, _) in m) for m in metas]
    ref = local[keys[0]][0] if keys else (packed if packed is not None else torch.zeros(0, dtype=torch.float64))
    dev = ref.device
    mine = sizes[rank]
    if packed is not None and with_vectors and packed.numel() >= mine and packed.is_contiguous():
        send = packed[:mine]
    else:
        send = torch.zeros(mine, dtype=torch.float64, device=dev)
        off = 0
        for i in keys:
            w, V = local[i][0], local[i][1]
            k = w.numel()
            send[off:off + k].copy_(w.reshape(-1))
            off += k
            if with_vectors:
                send[off:off + V.numel()].copy_(V.reshape(-1))
                off += V.numel()
    recv: List[Optional[torch.Tensor]] = [None] * world
    if not multi:
        recv[0] = send
    elif dst is None:
        pad = max(max(sizes), 1)
        sp = send if send.numel() == pad else torch.cat([send, torch.zeros(pad - send.numel(), dtype=send.dtype, device=dev)])
        bufs = [torch.empty(pad, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(bufs, sp, group=group)
        recv = [bufs[r][:sizes[r]] for r in range(world)]
    else:
        gdst = dist.get_global_rank(group, dst) if group is not None else dst
        if rank == dst:
            reqs = []
            for r in range(world):
                if r == dst:
                    recv[r] = send
                elif sizes[r]:
                    recv[r] = torch.empty(sizes[r], dtype=torch.float64, device=dev)
                    src = dist.get_global_rank(group, r) if group is not None else r
                    reqs.append(dist.irecv(recv[r], src=src, group=group))
                else:
                    recv[r] = torch.zeros(0, dtype=torch.float64, device=dev)
            for q in reqs:
                q.wait()
        elif mine:
            dist.send(send, dst=gdst, group=group)
    if dst is not None and rank != dst:
        if stats is not None:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            stats.update(bytes=0, seconds=_time.perf_counter() - t_start)
        return None
    out: List[Optional[tuple]] = [None] * n_problems
    nbytes = 0
    for r, m in enumerate(metas):
        buf = own_host if (r == rank and own_host is not None and own_host.size == sizes[r]) else _to_host(recv[r])
        nbytes += buf.nbytes
        off = 0
        for (i, k, n, cnt, rc) in m:
            w = buf[off:off + k].copy()
            off += k
            v = None
            if with_vectors:
                v = buf[off:off + k * n].reshape(k, n)      # a view of the (pinned) landing buffer: no second host copy
                off += k * n
            out[i] = (w, v, cnt, r, rc)
    if stats is not None:
        stats.update(bytes=int(nbytes), seconds=_time.perf_counter() - t_start)
    missing = [i for i, o in enumerate(out) if o is None]
    if missing:
        raise RuntimeError("sweep gather: no rank reported problems %s" % missing[:8])
    return out


# ---------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------
class _Lane:
    """One of the eigenproblems a GPU works on concurrently: its own CUDA stream, solver (front storage), material
    array and value arrays.  Everything that depends on the mesh only is shared through the _MeshContext."""

    def __init__(self, ctx, stream, workspace=None):
        from . import solver as sv
        self.ctx, self.stream = ctx, stream
        prob = ctx.prob
        with torch.cuda.stream(stream):
            self.solver = sv.Solver(prob.pattern, ctx.symbolic)
            if len(ctx.streams) > 1:
                # several lanes share the GPU: plain stream order inside a solve (measured +12 % problems/s at 8 lanes)
                from . import _lib
                _lib.check(_lib.lib().ws_solver_set_overlap(self.solver.handle, 0))
            self.k2n2 = torch.empty_like(prob.k2n2)
            nnz = prob.pattern.nnz
            self.M = torch.empty(nnz, dtype=torch.float64, device=prob.dev)
            self.B = torch.empty(nnz, dtype=torch.float64, device=prob.dev)
        import ctypes as C
        from . import _lib
        self.own_workspace = workspace is None
        if workspace is None:
            workspace = C.c_void_p()
            _lib.check(_lib.lib().ws_eig_workspace_create(C.byref(workspace)))
        self.workspace = workspace

    def close(self):
        from . import _lib
        self.stream.synchronize()
        if self.own_workspace and self.workspace is not None and self.workspace.value:
            _lib.lib().ws_eig_workspace_destroy(self.workspace)
        self.workspace = None
        self.solver.close()


class _MeshContext:
    """Everything that depends on the mesh only, resident on this rank's GPU: arrays, pattern, incidence lists,
    symbolic analysis (shared by the lanes), plus the lanes themselves."""

    def __init__(self, mesh, kind, IOR_dict, k, device, leaf_elems, streams=None, workspaces=None):
        from . import fe_solver as fs, solver as sv
        self.kind = kind
        self.mesh = mesh
        if kind == "vector":
            self.prob = fs.VectorProblem(mesh, IOR_dict, k, device, build_pattern=False)
            fcs = self.prob.Ntt
        else:
            self.prob = fs.ScalarProblem(mesh, IOR_dict, k, device, build_pattern=False)
            fcs = 0
        self.prob.build_pattern()
        self.symbolic = sv.Symbolic.on_device(self.prob.dofs, self.prob.xy, self.prob.N, fcs,
                                              leaf_elems or sv.DEFAULT_LEAF_ELEMS)
        self.lanes = []
        self.streams = list(streams or [])
        if streams is None:
            self.solver = sv.Solver(self.prob.pattern, self.symbolic)      # single-lane use (tools, tests)
        else:
            self.solver = None
            if kind == "vector":
                # the T / T^T tables hang off the shared pattern and are built on first use: build them before the
                # lanes (other streams) read them
                self.prob.apply_T(torch.zeros(self.prob.N, dtype=torch.float64, device=self.prob.dev))
            cur = torch.cuda.current_stream(self.prob.dev)
            for li, st in enumerate(streams):
                st.wait_stream(cur)
                self.lanes.append(_Lane(self, st, workspaces[li] if workspaces else None))

    def close(self):
        for ln in self.lanes:
            ln.close()
        if self.solver is not None:
            self.solver.close()
        self.prob.pattern.close()
        self.symbolic.close()


def default_lanes(n_unknowns: int) -> int:
    """Problems a B200 works on concurrently: an eigensolve of ~100k unknowns is a chain of ~4000 small dependent
    kernels that fills a fraction of the GPU (4 streams finish 3.2 solves in the time of one); from ~1M unknowns
    on a single problem saturates it."""
    import os
    if os.environ.get("WS_SWEEP_LANES"):          # experiments (profiles/r02c_small_problem_top_levels.txt)
        return max(1, int(os.environ["WS_SWEEP_LANES"]))
    if n_unknowns >= 1_000_000:
        return 1
    if n_unknowns >= 400_000:
        return 2
    return 8


def solve_sweep(meshes: Sequence, problems: Sequence[SweepProblem], Nmax: int = 10, kind: str = "vector",
                group=None, device=None, leaf_elems=None, with_vectors: bool = True, gather: bool = True,
                dst: Optional[int] = None, lanes: Optional[int] = None, gather_stats: Optional[dict] = None):
    """Solve every problem of the sweep, this rank taking its share (``partition``), and gather.

    ``meshes``: mesh objects (vector sweeps: already passed through ``get_unique_edges``), the same
    list on every rank.  Returns a list of ``SweepResult`` in problem order (on every rank, or on
    ``dst`` only), or this rank's own results when ``gather=False``.  A problem that fails (no convergence,
    inaccurate factorisation at its shift, ...) is reported through ``SweepResult.rc`` / ``info["error"]``
    and does not stop the others.

    ``lanes``: problems solved concurrently on this GPU (default: ``default_lanes``).  The lanes of a mesh share
    its pattern, incidence lists and symbolic analysis; each has its own CUDA stream, solver storage and value
    arrays.  ONE host thread drives them in lockstep: re-assembly and factorisation of the next wavelength of
    every lane are enqueued back to back (``ws_solver_factor_begin`` / ``_end``), then the eigen-iterations run
    interleaved through ``ws_eig_solve_batch``, so the kernels of the lanes overlap on the device.  (Round 1 drove
    the lanes from one host thread each: the threads serialised on the driver and 4 lanes were no faster than
    one.)"""
    import ctypes as C
    from . import _lib, device as _dev, fe_solver as fs
    assert kind in ("vector", "scalar")
    dist = _dist()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    rank = dist.get_rank(group) if multi else 0
    world = dist.get_world_size(group) if multi else 1
    dev = _dev.require_cuda(device)
    L = _lib.lib()

    def n_unknowns(mesh):
        # the same rule the problem classes apply (fe_solver.VectorProblem / ScalarProblem.N): the packed
        # send-buffer offsets and the gather metadata must agree
        if kind == "vector":
            return len(mesh.cells[0].data) + len(mesh.points)
        conn = np.asarray(mesh.cells[1].data)
        used = [conn[np.asarray(sel[1])] if sel[1] is not None else conn for sel in mesh.cell_sets.values()]
        used = [u for u in used if u.size]
        return int(max(int(u.max()) for u in used)) + 1 if used else 0

    sizes = [n_unknowns(m) for m in meshes]
    table = partition([p.mesh_id for p in problems], [estimate_cost(sizes[p.mesh_id]) for p in problems], world)
    mine = table[rank]
    if lanes is None:
        lanes = default_lanes(max([sizes[problems[i].mesh_id] for i in mine], default=0))
    lanes = max(1, min(int(lanes), max(1, len(mine))))
    local, infos = {}, {}
    # one flat send buffer per rank: [w_0 | V_0 | w_1 | V_1 ...]; the eigensolver writes into its views
    total = sum(Nmax * (1 + sizes[problems[i].mesh_id]) for i in mine)
    with torch.cuda.device(dev):
        main_stream = torch.cuda.current_stream(dev)
        sendbuf = torch.zeros(max(total, 1), dtype=torch.float64, device=dev)
        offsets, off = {}, 0
        for i in mine:
            offsets[i] = off
            off += Nmax * (1 + sizes[problems[i].mesh_id])
        streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
        # one eigensolver work space per lane for the whole sweep: its pinned staging blocks survive the change of
        # mesh (the basis and the step graphs are rebuilt when the sizes / handles change)
        workspaces = []
        for _ in range(lanes):
            h_ = C.c_void_p()
            _lib.check(L.ws_eig_workspace_create(C.byref(h_)))
            workspaces.append(h_)
        # the problems of one mesh, in order; `lanes` of them are in flight at a time
        groups: List[List[int]] = []
        for i in mine:
            if groups and problems[groups[-1][-1]].mesh_id == problems[i].mesh_id:
                groups[-1].append(i)
            else:
                groups.append([i])

        def fail(i, wv, Vv, e):
            rc = getattr(e, "rc", 0) or _lib.ERR_CUDA
            if rc != _lib.ERR_NOCONV:
                wv.fill_(float("nan"))
                Vv.fill_(float("nan"))
            return rc

        import os as _os, time as _time
        _trace = bool(_os.environ.get("WS_SWEEP_TRACE"))
        _tm = {"context": 0.0, "assemble+factor_begin": 0.0, "factor_end": 0.0, "eig_batch": 0.0, "post": 0.0, "close": 0.0}

        def _tick(name, t0):
            if _trace:
                _tm[name] += _time.perf_counter() - t0
            return _time.perf_counter()

        try:
            for grp in groups:
                pr0 = problems[grp[0]]
                _t = _time.perf_counter()
                ctx = _MeshContext(meshes[pr0.mesh_id], kind, pr0.IOR_dict, 2 * np.pi / pr0.wl, dev, leaf_elems,
                                   streams=streams[:min(lanes, len(grp))], workspaces=workspaces)
                prob = ctx.prob
                N = prob.N
                if _trace:
                    torch.cuda.synchronize(dev)
                _t = _tick("context", _t)
                try:
                    for r0 in range(0, len(grp), len(ctx.lanes)):
                        batch = grp[r0:r0 + len(ctx.lanes)]
                        _t = _time.perf_counter()
                        act = []                                   # (problem index, lane, k, sigma, views)
                        # ---- phase 1: new materials, re-assembly (K11) and factorisation of every lane, all enqueued
                        for i, ln in zip(batch, ctx.lanes):
                            pr = problems[i]
                            k = 2 * np.pi / pr.wl
                            o = offsets[i]
                            wv = sendbuf[o:o + Nmax]
                            Vv = sendbuf[o + Nmax:o + Nmax + Nmax * N].view(Nmax, N)
                            nmax_ior = max(pr.IOR_dict.values())
                            sigma = float(np.power(k * (nmax_ior if pr.target_neff is None else pr.target_neff), 2))
                            sp = ln.stream.cuda_stream
                            try:
                                with torch.cuda.stream(ln.stream):
                                    ln.k2n2.copy_(torch.from_numpy(fs._k2n2_host(ctx.mesh, pr.IOR_dict, k)), non_blocking=True)
                                    if kind == "vector":
                                        _lib.check(L.ws_assemble_vector_qd(prob.pattern.handle, _dev.ptr(prob.xy), _dev.ptr(prob.tris),
                                                                           _dev.ptr(ln.k2n2), prob.Ntt, sigma, _dev.ptr(ln.M),
                                                                           _dev.ptr(ln.B), _lib.ASM_FAST, dev.index, sp))
                                        _lib.check(L.ws_solver_factor_begin(ln.solver.handle, _dev.ptr(ln.M), None, 0.0, dev.index, sp))
                                    else:
                                        _lib.check(L.ws_assemble_scalar_p2(prob.pattern.handle, _dev.ptr(prob.xy), _dev.ptr(prob.dofs),
                                                                           _dev.ptr(ln.k2n2), _dev.ptr(ln.M), _dev.ptr(ln.B),
                                                                           _lib.ASM_FAST, dev.index, sp))
                                        _lib.check(L.ws_solver_factor_begin(ln.solver.handle, _dev.ptr(ln.M), _dev.ptr(ln.B), sigma,
                                                                            dev.index, sp))
                                act.append((i, ln, k, sigma, wv, Vv))
                            except _lib.WavesolveB200Error as e:
                                local[i] = (wv, Vv, 0, fail(i, wv, Vv, e))
                                infos[i] = {"error": str(e)}
                        _t = _tick("assemble+factor_begin", _t)
                        # ---- phase 2: read the pivot statistics / accuracy probe of every lane
                        ready = []
                        for (i, ln, k, sigma, wv, Vv) in act:
                            try:
                                _lib.check(L.ws_solver_factor_end(ln.solver.handle, dev.index, ln.stream.cuda_stream))
                                ready.append((i, ln, k, sigma, wv, Vv))
                            except _lib.WavesolveB200Error as e:
                                local[i] = (wv, Vv, 0, fail(i, wv, Vv, e))
                                infos[i] = {"error": str(e)}
                        _t = _tick("factor_end", _t)
                        if not ready:
                            continue
                        # ---- phase 3: the eigen-iterations of the lanes, interleaved by one host thread
                        jobs = (_lib.EigJob * len(ready))()
                        for J, (i, ln, k, sigma, wv, Vv) in zip(jobs, ready):
                            J.pattern, J.solver, J.B_vals = prob.pattern.handle.value, ln.solver.handle.value, _dev.ptr(ln.B)
                            J.sigma, J.kind, J.nev, J.ncv, J.maxit, J.tol, J.seed = sigma, (1 if kind == "vector" else 0), Nmax, 0, 0, 0.0, 0
                            J.edges = _dev.ptr(prob.edges) if kind == "vector" else None
                            J.xy = _dev.ptr(prob.xy) if kind == "vector" else None
                            J.Ntt = prob.Ntt if kind == "vector" else 0
                            J.w, J.V, J.stream, J.resid = _dev.ptr(wv), _dev.ptr(Vv), ln.stream.cuda_stream, None
                            J.workspace = ln.workspace.value
                        _lib.check(L.ws_eig_solve_batch(C.addressof(jobs), len(ready), dev.index))
                        _t = _tick("eig_batch", _t)
                        for J, (i, ln, k, sigma, wv, Vv) in zip(jobs, ready):
                            pr = problems[i]
                            rc = int(J.rc)
                            info = dict(nconv=int(J.info[0]), n_op=int(J.info[1]), restarts=int(J.info[2]), ncomplex=int(J.info[3]),
                                        N=N, sigma=sigma, factor=ln.solver.stats())
                            cnt = 0
                            if rc not in (0, _lib.ERR_NOCONV):
                                wv.fill_(float("nan"))
                                Vv.fill_(float("nan"))
                            if rc:
                                info["error"] = J.error.decode(errors="replace")
                            if rc in (0, _lib.ERR_NOCONV):
                                cnt = fs._count_modes(wv.cpu().numpy(), k, pr.IOR_dict,
                                                      always=(kind == "vector" and pr.target_neff is not None))
                            local[i] = (wv, Vv, cnt, rc)
                            infos[i] = info
                        _t = _tick("post", _t)
                finally:
                    _t = _time.perf_counter()
                    for st_ in streams:
                        st_.synchronize()
                    ctx.close()
                    _t = _tick("close", _t)
            if _trace:
                import sys as _sys
                print("[solve_sweep] rank %d lanes %d problems %d: %s" % (rank, lanes, len(mine), {k_: round(v, 3) for k_, v in _tm.items()}),
                      file=_sys.stderr, flush=True)
        finally:
            for st_ in streams:
                st_.synchronize()
            for h_ in workspaces:
                L.ws_eig_workspace_destroy(h_)
        for st_ in streams:
            st_.synchronize()
            main_stream.wait_stream(st_)
        local = {i: local[i] for i in mine}      # send-buffer order
        torch.cuda.synchronize(dev)
        if not gather:
            return [SweepResult(i, local[i][0].cpu().numpy(), local[i][1].cpu().numpy() if with_vectors else None,
                                local[i][2], rank, infos[i], local[i][3]) for i in mine]
        got = gather_eigenpairs(local, len(problems), group=group, with_vectors=with_vectors, dst=dst,
                                packed=sendbuf, stats=gather_stats)
    if got is None:
        return None
    return [SweepResult(i, w, v, cnt, r, infos.get(i, {}), rc) for i, (w, v, cnt, r, rc) in enumerate(got)]
