// K8, symbolic phase (host, C++): element-based geometric nested dissection and the front
// structure of the multifrontal factorisation.
//
//  1. k-d bisection of the element centroids into a complete binary tree of 2^depth leaves
//     (median split along the longer bounding-box axis).
//  2. every unknown is owned by the lowest common ancestor of the leaves of its elements
//     (longest common prefix of the smallest and largest leaf index).
//  3. elimination order: level-major bottom-up; inside a front the unknowns with old index
//     >= first_class_start come first (vector problem: nodes before edges, which keeps the
//     pivot blocks of the quasi-definite matrix non-singular), then ascending old index.
//  4. front(t) = own(t) ++ boundary(t), boundary(t) = unknowns touched by the elements of
//     subtree(t) that are owned by proper ancestors, ascending new index.
//
// The reference leaves all of this to SuperLU's COLAMD ordering inside scipy
// (fe_solver.py:328, :398); having the mesh coordinates makes a geometric dissection both
// cheaper and better (nnz(L)/N ~ 70-90 against ~115 with COLAMD on the same matrices).
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <thread>

#include "ws_solver.cuh"

namespace ws {

namespace {

struct Pt {
    double x, y;
    int32_t id;
};

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void kd_rec(Pt* a, int64_t lo, int64_t hi, int level, int depth, int64_t idx, int64_t* leaf_lo, int par_levels) {
    if (level == depth) {
        leaf_lo[idx] = lo;
        return;
    }
    const int64_t mid = (lo + hi) / 2;
    if (hi - lo > 1) {
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int64_t i = lo; i < hi; ++i) {
            x0 = std::min(x0, a[i].x); x1 = std::max(x1, a[i].x);
            y0 = std::min(y0, a[i].y); y1 = std::max(y1, a[i].y);
        }
        if (x1 - x0 >= y1 - y0)
            std::nth_element(a + lo, a + mid, a + hi, [](const Pt& u, const Pt& v) { return u.x < v.x || (u.x == v.x && u.id < v.id); });
        else
            std::nth_element(a + lo, a + mid, a + hi, [](const Pt& u, const Pt& v) { return u.y < v.y || (u.y == v.y && u.id < v.id); });
    }
    if (level < par_levels) {
        std::thread th(kd_rec, a, lo, mid, level + 1, depth, 2 * idx, leaf_lo, par_levels);
        kd_rec(a, mid, hi, level + 1, depth, 2 * idx + 1, leaf_lo, par_levels);
        th.join();
    } else {
        kd_rec(a, lo, mid, level + 1, depth, 2 * idx, leaf_lo, par_levels);
        kd_rec(a, mid, hi, level + 1, depth, 2 * idx + 1, leaf_lo, par_levels);
    }
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
    if (n <= 0) return;
    nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, n));
    if (nthreads == 1) {
        for (int64_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=]() {
            for (int64_t i = t; i < n; i += nthreads) f(i);
        });
    for (auto& x : th) x.join();
}

}  // namespace

void analyse_host(Symbolic& S, const int32_t* dofs, int64_t Ne, int64_t N, const double* xy, int64_t Np,
                  int vert_col, int64_t vert_offset, int64_t cls_start, int leaf_elems) {
    WS_REQUIRE(Ne > 0 && N > 0, "analyse: empty problem");
    WS_REQUIRE(leaf_elems >= 1, "leaf_elems >= 1");
    int depth = 0;
    while ((Ne >> depth) > leaf_elems && depth < 24) ++depth;
    S.N = N;
    S.Ne = Ne;
    S.depth = depth;
    const int64_t nleaf = (int64_t)1 << depth;
    const int64_t nfronts = (int64_t)1 << (depth + 1);
    S.nfronts = nfronts;
    int hw = (int)std::thread::hardware_concurrency();
    const int nthreads = std::max(1, std::min(hw ? hw : 4, 32));
    int par_levels = 0;
    while ((1 << par_levels) < nthreads && par_levels < depth) ++par_levels;

    // ---- 1. k-d bisection
    double t0 = now();
    std::vector<Pt> pts(Ne);
    {
        // element centroids from the three vertex ids stored at dofs[e*6 + vert_col ..] - vert_offset
        std::vector<int> bad(nthreads, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                const int64_t e0 = Ne * t / nthreads, e1 = Ne * (t + 1) / nthreads;
                for (int64_t e = e0; e < e1; ++e) {
                    double cx = 0.0, cy = 0.0;
                    for (int a = 0; a < 3; ++a) {
                        const int64_t v = (int64_t)dofs[e * 6 + vert_col + a] - vert_offset;
                        if (v < 0 || v >= Np) { bad[t] = 1; continue; }
                        cx += xy[2 * v];
                        cy += xy[2 * v + 1];
                    }
                    pts[e] = Pt{cx / 3.0, cy / 3.0, (int32_t)e};
                }
            });
        for (auto& x : th) x.join();
        for (int b : bad) WS_REQUIRE(!b, "analyse: vertex id out of range");
    }
    std::vector<int64_t> leaf_lo(nleaf + 1, 0);
    kd_rec(pts.data(), 0, Ne, 0, depth, 0, leaf_lo.data(), par_levels);
    leaf_lo[nleaf] = Ne;
    std::vector<int32_t> leaf_of(Ne);
    for (int64_t j = 0; j < nleaf; ++j)
        for (int64_t i = leaf_lo[j]; i < leaf_lo[j + 1]; ++i) leaf_of[pts[i].id] = (int32_t)j;
    S.t_kd = now() - t0;

    // ---- 2. owners (parallel over elements; min / max leaf per unknown with relaxed CAS)
    t0 = now();
    std::vector<int32_t> lo(N, INT_MAX), hi(N, -1);
    {
        int32_t* plo = lo.data();
        int32_t* phi = hi.data();
        std::vector<int> bad(nthreads, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                const int64_t e0 = Ne * t / nthreads, e1 = Ne * (t + 1) / nthreads;
                for (int64_t e = e0; e < e1; ++e) {
                    const int32_t lf = leaf_of[e];
                    for (int a = 0; a < 6; ++a) {
                        const int32_t d = dofs[e * 6 + a];
                        if (d < 0 || d >= N) { bad[t] = 1; continue; }
                        int32_t cur = __atomic_load_n(plo + d, __ATOMIC_RELAXED);
                        while (lf < cur && !__atomic_compare_exchange_n(plo + d, &cur, lf, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                        cur = __atomic_load_n(phi + d, __ATOMIC_RELAXED);
                        while (lf > cur && !__atomic_compare_exchange_n(phi + d, &cur, lf, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
                    }
                }
            });
        for (auto& x : th) x.join();
        for (int b : bad) WS_REQUIRE(!b, "analyse: unknown id out of range");
    }
    // ---- 3. elimination order: level-major bottom-up, class, old index (stable parallel counting sort)
    std::vector<int64_t> levelbase(depth + 2, 0);
    for (int l = depth; l >= 1; --l) levelbase[l - 1] = levelbase[l] + ((int64_t)1 << l);
    auto prank = [&](int64_t nd) { int l = 63 - __builtin_clzll((unsigned long long)nd); return levelbase[l] + nd - ((int64_t)1 << l); };
    const int64_t nbuckets = 2 * nfronts;
    std::vector<int32_t> bucket(N);
    std::vector<std::vector<int64_t>> hist(nthreads);
    std::vector<int64_t> missing(nthreads, -1);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                hist[t].assign(nbuckets, 0);
                const int64_t d0 = N * t / nthreads, d1 = N * (t + 1) / nthreads;
                for (int64_t d = d0; d < d1; ++d) {
                    if (hi[d] < 0) { if (missing[t] < 0) missing[t] = d; bucket[d] = 0; continue; }
                    const uint32_t x = (uint32_t)(lo[d] ^ hi[d]);
                    const int hb = x ? 32 - __builtin_clz(x) : 0;
                    const int64_t nd = (((int64_t)1 << depth) + lo[d]) >> hb;
                    const int32_t b = (int32_t)(2 * prank(nd) + (d >= cls_start ? 0 : 1));
                    bucket[d] = b;
                    hist[t][b]++;
                }
            });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < nthreads; ++t)
        if (missing[t] >= 0)
            throw Error(WS_ERR_NUMERIC, "matrix is singular: unknown " + std::to_string(missing[t]) +
                                            " belongs to no element (empty row)");
    // exclusive prefix over (bucket, thread)
    std::vector<int64_t> bstart(nbuckets + 1, 0);
    {
        int64_t run = 0;
        for (int64_t b = 0; b < nbuckets; ++b) {
            bstart[b] = run;
            for (int t = 0; t < nthreads; ++t) {
                const int64_t c = hist[t][b];
                hist[t][b] = run;
                run += c;
            }
        }
        bstart[nbuckets] = run;
    }
    S.p.assign(nfronts, 0);
    S.q.assign(nfronts, 0);
    S.start.assign(nfronts, 0);
    for (int64_t nd = 1; nd < nfronts; ++nd) {
        const int64_t r = prank(nd);
        S.start[nd] = bstart[2 * r];
        S.p[nd] = (int32_t)(bstart[2 * r + 2] - bstart[2 * r]);
    }
    S.perm.resize(N);
    S.iperm.resize(N);
    S.node_of_new.resize(N);
    {
        // bucket -> front id (inverse of prank)
        std::vector<int32_t> node_of_bucket(nbuckets);
        for (int64_t nd = 1; nd < nfronts; ++nd) {
            const int64_t r = prank(nd);
            node_of_bucket[2 * r] = node_of_bucket[2 * r + 1] = (int32_t)nd;
        }
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                const int64_t d0 = N * t / nthreads, d1 = N * (t + 1) / nthreads;
                int64_t* cur = hist[t].data();
                for (int64_t d = d0; d < d1; ++d) {
                    const int32_t b = bucket[d];
                    const int64_t nw = cur[b]++;
                    S.iperm[d] = (int32_t)nw;
                    S.perm[nw] = (int32_t)d;
                    S.node_of_new[nw] = node_of_bucket[b];
                }
            });
        for (auto& x : th) x.join();
    }
    S.t_order = now() - t0;

    // ---- 4. boundaries
    t0 = now();
    std::vector<std::vector<int32_t>> B(nfronts);
    parallel_for(nleaf, nthreads, [&](int64_t j) {
        const int64_t t = nleaf + j;
        std::vector<int32_t> v;
        v.reserve((leaf_lo[j + 1] - leaf_lo[j]) * 6);
        for (int64_t i = leaf_lo[j]; i < leaf_lo[j + 1]; ++i) {
            const int32_t* d = dofs + (int64_t)pts[i].id * 6;
            for (int a = 0; a < 6; ++a) v.push_back(S.iperm[d[a]]);
        }
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        const int64_t own_end = S.start[t] + S.p[t];
        auto it = std::lower_bound(v.begin(), v.end(), (int32_t)own_end);
        B[t].assign(it, v.end());
    });
    for (int l = depth - 1; l >= 0; --l) {
        const int64_t base = (int64_t)1 << l;
        parallel_for(base, nthreads, [&](int64_t j) {
            const int64_t t = base + j;
            const auto& a = B[2 * t];
            const auto& b = B[2 * t + 1];
            std::vector<int32_t> u(a.size() + b.size());
            auto e = std::set_union(a.begin(), a.end(), b.begin(), b.end(), u.begin());
            u.resize(e - u.begin());
            const int64_t own_end = S.start[t] + S.p[t];
            auto it = std::lower_bound(u.begin(), u.end(), (int32_t)own_end);
            // every own unknown of t is shared by both children, hence in the union
            if ((int64_t)(it - u.begin()) != S.p[t] || (S.p[t] && u.front() != (int32_t)S.start[t])) {
                B[t].clear();
                B[t].push_back(-1);   // flagged below (cannot throw inside a worker thread)
                return;
            }
            B[t].assign(it, u.end());
        });
        for (int64_t j = 0; j < base; ++j)
            if (B[base + j].size() == 1 && B[base + j][0] == -1)
                throw Error(WS_ERR_STATE, "internal: separator ownership inconsistent");
    }
    S.bndoff.assign(nfronts + 1, 0);
    for (int64_t t = 1; t < nfronts; ++t) {
        S.q[t] = (int32_t)B[t].size();
        S.bndoff[t + 1] = S.bndoff[t] + S.q[t];
    }
    S.bndoff[1] = 0;
    for (int64_t t = 1; t < nfronts; ++t) S.bndoff[t + 1] = S.bndoff[t] + S.q[t];
    S.bnd.resize(S.bndoff[nfronts]);
    S.rel.assign(S.bndoff[nfronts], -1);
    parallel_for(nfronts - 1, nthreads, [&](int64_t i) {
        const int64_t t = i + 1;
        std::copy(B[t].begin(), B[t].end(), S.bnd.begin() + S.bndoff[t]);
        if (t == 1) return;
        const int64_t par = t / 2;
        const int64_t ps = S.start[par], pe = ps + S.p[par];
        const auto& pb = B[par];
        size_t k = 0;
        for (size_t j = 0; j < B[t].size(); ++j) {
            const int32_t g = B[t][j];
            int32_t r;
            if (g < pe) {
                r = (int32_t)(g - ps);
            } else {
                while (k < pb.size() && pb[k] < g) ++k;
                r = (int32_t)(S.p[par] + k);
            }
            S.rel[S.bndoff[t] + j] = r;
        }
    });
    // ---- offsets and statistics
    S.Loff.assign(nfronts, 0);
    S.Uoff.assign(nfronts, 0);
    S.woff.assign(nfronts, 0);
    int64_t Lt = 0, Ut = 0, Wt = 0;
    S.nnzL = 0;
    S.flops = 0;
    S.maxfront = 0;
    S.maxp = 0;
    for (int64_t t = 1; t < nfronts; ++t) {
        const int64_t p = S.p[t], q = S.q[t], m = p + q;
        S.Loff[t] = Lt; Lt += (m * p + 1) & ~(int64_t)1;
        S.Uoff[t] = Ut; Ut += (q * q + 1) & ~(int64_t)1;
        S.woff[t] = Wt; Wt += (m + 1) & ~(int64_t)1;
        S.nnzL += m * p;
        S.flops += (double)p * p * p / 3.0 + (double)p * p * q + (double)p * q * q;
        S.maxfront = std::max<int>(S.maxfront, (int)m);
        S.maxp = std::max<int>(S.maxp, (int)p);
    }
    S.Ltotal = Lt; S.Utotal = Ut; S.Wtotal = Wt;
    S.t_bnd = now() - t0;
}

}  // namespace ws

using namespace ws;

extern "C" {

int ws_symbolic_create(const int32_t* h_dofs, int64_t Ne, int64_t N, const double* h_xy, int64_t Np,
                       int vert_col, int64_t vert_offset, int64_t first_class_start, int leaf_elems,
                       ws_symbolic** out) {
    WS_API_BEGIN
    WS_REQUIRE(out, "out == NULL");
    *out = nullptr;
    WS_REQUIRE(h_dofs && h_xy, "null pointer");
    WS_REQUIRE(vert_col == 0 || vert_col == 3, "vert_col must be 0 or 3");
    WS_REQUIRE(N < ((int64_t)1 << 31), "N must fit int32");
    auto* s = new ws_symbolic();
    try {
        analyse_host(s->S, h_dofs, Ne, N, h_xy, Np, vert_col, vert_offset, first_class_start, leaf_elems);
    } catch (...) {
        delete s;
        throw;
    }
    *out = s;
    WS_API_END
}

int ws_symbolic_query(const ws_symbolic* s, int64_t* h_info, double* h_dinfo) {
    WS_API_BEGIN
    WS_REQUIRE(s, "symbolic == NULL");
    const Symbolic& S = s->S;
    if (h_info) {
        h_info[0] = S.N; h_info[1] = S.depth; h_info[2] = S.nfronts; h_info[3] = S.nnzL;
        h_info[4] = S.Utotal; h_info[5] = S.maxfront; h_info[6] = S.maxp;
        h_info[7] = s->on_device ? s->sumq : (int64_t)S.bnd.size();
    }
    if (h_dinfo) {
        h_dinfo[0] = S.flops; h_dinfo[1] = S.t_kd; h_dinfo[2] = S.t_order; h_dinfo[3] = S.t_bnd;
    }
    WS_API_END
}

int ws_symbolic_export(const ws_symbolic* s, int32_t* h_perm, int32_t* h_p, int32_t* h_q,
                       int64_t* h_start, int64_t* h_bndoff, int32_t* h_bnd, int32_t* h_rel) {
    WS_API_BEGIN
    WS_REQUIRE(s, "symbolic == NULL");
    const Symbolic& S = s->S;
    if (s->on_device) {
        DeviceGuard g(s->device);
        if (h_perm) WS_CUDA(cudaMemcpy(h_perm, s->d_perm.get(), S.N * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (h_bnd && s->sumq) WS_CUDA(cudaMemcpy(h_bnd, s->d_bnd.get(), s->sumq * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (h_rel && s->sumq) WS_CUDA(cudaMemcpy(h_rel, s->d_rel.get(), s->sumq * sizeof(int32_t), cudaMemcpyDeviceToHost));
        h_perm = nullptr;
        h_bnd = h_rel = nullptr;
    }
    if (h_perm) std::copy(S.perm.begin(), S.perm.end(), h_perm);
    if (h_p) std::copy(S.p.begin(), S.p.end(), h_p);
    if (h_q) std::copy(S.q.begin(), S.q.end(), h_q);
    if (h_start) std::copy(S.start.begin(), S.start.end(), h_start);
    if (h_bndoff) std::copy(S.bndoff.begin(), S.bndoff.end(), h_bndoff);
    if (h_bnd) std::copy(S.bnd.begin(), S.bnd.end(), h_bnd);
    if (h_rel) std::copy(S.rel.begin(), S.rel.end(), h_rel);
    WS_API_END
}

int ws_symbolic_destroy(ws_symbolic* s) {
    WS_API_BEGIN
    if (s && s->on_device) {
        DeviceGuard g(s->device);
        delete s;
    } else {
        delete s;
    }
    WS_API_END
}

}  // extern "C"
