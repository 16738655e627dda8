// K8, numeric phase: multifrontal LDL^T on the device + selective inversion + solve sweeps.
//
// Everything is expressed as batched *tile tasks*: the host (ws_solver_create) walks the
// front tree once and emits task records for generic kernels --
//     k_gemm_tiles   C(64xBN tile) (+)= (-) A * op(B)        fp64 tensor cores (DMMA m8n8k4)
//     k_panel        NBxNB diagonal block LDL^T + row-block TRSM
//     k_diag_inverse unit-lower NBxNB inverse
//     k_extend_add   child update matrix -> parent front
//     k_fwd_gemv / k_bwd_gemvT   the two solve sweeps as dense mat-vecs
// -- and the numeric phase is a fixed sequence of launches over slices of those task arrays,
// level by level from the leaves to the root (left-looking inside a front: every block column
// is updated once with a rank-k0 GEMM, the Schur complement with one rank-p SYRK, so each
// output tile is read and written once).  All fronts are resident at once (HBM is 180 GB;
// the 1M-element problem needs ~12 GB), so there is no stack management and the update
// matrices of all children are available to the extend-add of a level in one launch.
//
// Replaces SuperLU (scipy splu / spsolve) behind fe_solver.py:328 and fe_solver.py:398.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

#include "ws_solver.cuh"

namespace ws {

struct Range {
    int64_t off = 0, cnt = 0;
};

struct ExtTask {
    int32_t child, ti, tj;
};

struct InvTask {
    double* S;       // diagonal block (k0,k0) of the solve panel: holds L_kk (strict lower) + d
    int32_t ld, nb;
};

struct LevelPlan {
    std::vector<Range> upd, pan, trail;   // per block column (upd: left-looking, trail: right-looking)
    Range syrk, ext_left, ext_right, fwd, bwd;
    int64_t first_front = 0, nfronts = 0;
    bool small = false;                   // every front fits the single-CTA solve kernels
    bool right_looking = false;
};

}  // namespace ws

struct ws_solver {
    int device = 0;
    int64_t N = 0, nnz = 0, nfronts = 0;
    int depth = 0;
    ws::Symbolic sym;    // host copy of the structure (small)
    // structure on device
    ws::DevBuf<int32_t> perm, iperm, node_of_new, fr_p, fr_q, bnd, rel;
    ws::DevBuf<int64_t> fr_start, fr_Loff, fr_Uoff, fr_woff, fr_bndoff, dest;
    // numeric storage
    ws::DevBuf<double> L, S, T, U, W, d, dinv, y, x, bperm;
    ws::DevBuf<unsigned long long> stats;   // [0] negative pivots [1] zero/NaN pivots [2] min |pivot| bits [3] max |pivot| bits
    // tasks
    ws::DevBuf<ws::GemmTask> gemm_tasks;
    ws::DevBuf<ws::PanelTask> panel_tasks;
    ws::DevBuf<ws::ExtTask> ext_tasks;
    ws::DevBuf<ws::InvTask> inv_tasks;
    ws::DevBuf<ws::GemvTask> gemv_tasks;
    ws::DevBuf<ws::GemvTTask> gemvt_tasks;
    std::vector<ws::LevelPlan> levels;      // index = tree level
    ws::Range inv0;
    std::vector<std::pair<ws::Range, ws::Range>> inv_stages;   // (step1, step2) per doubling stage
    ws::Range w21;
    bool factored = false;
    int64_t h_neg = 0, h_bad = 0;
    double h_minpiv = 0, h_maxpiv = 0;
    int64_t bytes = 0;
};

namespace ws {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int KC = 16;
constexpr int GSTAGES = 3;

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;                       // src-size 0 -> the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int BN>
constexpr size_t gemm_smem_bytes() {
    return (size_t)GSTAGES * (KC * (GT + 4) + KC * (BN + 4) + KC) * sizeof(double);
}

// One CTA = one GemmTask = one (<=64) x (<=BN) output tile.  3-stage cp.async pipeline over
// k-chunks of 16, fp64 tensor-core MMA (mma.sync m8n8k4) from shared memory, 2x2 warps.
template <int BN>
__global__ void __launch_bounds__(128) k_gemm_tiles(const GemmTask* __restrict__ tasks) {
    const GemmTask t = tasks[blockIdx.x];
    extern __shared__ double gsm[];
    constexpr int AS = KC * (GT + 4), BS = KC * (BN + 4);
    double* As = gsm;
    double* Bs = gsm + GSTAGES * AS;
    double* Ds = Bs + GSTAGES * BS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;
    constexpr int NBLK = BN / 16;
    constexpr int WN = BN / 2;
    double acc[4][NBLK][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NBLK; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    const bool bn = (t.flags & GF_BN) != 0;
    const int nk = (t.k + KC - 1) / KC;

    auto load_stage = [&](int stage, int kc) {
        double* as = As + stage * AS;
        double* bs = Bs + stage * BS;
        {
            const int i = tid & 63;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int kk = (tid >> 6) + 2 * r;
                const bool ok = i < t.m && kc + kk < t.k;
                cp_async8(as + kk * (GT + 4) + i, ok ? t.A + i + (size_t)(kc + kk) * t.lda : t.A, ok);
            }
        }
        if (!bn) {
#pragma unroll
            for (int r = 0; r < KC * BN / 128; ++r) {
                const int idx = tid + 128 * r;
                const int kk = idx / BN, jj = idx % BN;
                const bool ok = jj < t.n && kc + kk < t.k;
                cp_async8(bs + kk * (BN + 4) + jj, ok ? t.B + jj + (size_t)(kc + kk) * t.ldb : t.B, ok);
            }
        } else {
#pragma unroll
            for (int r = 0; r < KC * BN / 128; ++r) {
                const int idx = tid + 128 * r;
                const int kk = idx % KC, jj = idx / KC;
                const bool ok = jj < t.n && kc + kk < t.k;
                cp_async8(bs + kk * (BN + 4) + jj, ok ? t.B + (kc + kk) + (size_t)jj * t.ldb : t.B, ok);
            }
        }
        if (tid < KC) {
            if (t.dscale) {
                const bool ok = kc + tid < t.k;
                cp_async8(Ds + stage * KC + tid, ok ? t.dscale + kc + tid : t.dscale, ok);
            } else {
                Ds[stage * KC + tid] = 1.0;
            }
        }
    };

#pragma unroll
    for (int s = 0; s < GSTAGES; ++s) {
        if (s < nk) load_stage(s, s * KC);
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<GSTAGES - 1>();
        __syncthreads();
        const int stage = it % GSTAGES;
        const double* as = As + stage * AS;
        const double* bs = Bs + stage * BS;
        const double* ds = Ds + stage * KC;
#pragma unroll
        for (int ks = 0; ks < KC / 4; ++ks) {
            double a[4], b[NBLK];
            const double dk = ds[ks * 4 + tq];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) a[mb] = as[(ks * 4 + tq) * (GT + 4) + wm * 32 + mb * 8 + g];
#pragma unroll
            for (int nb = 0; nb < NBLK; ++nb) b[nb] = bs[(ks * 4 + tq) * (BN + 4) + wn * WN + nb * 8 + g] * dk;
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < NBLK; ++nb) dmma8x8x4(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
        }
        __syncthreads();
        if (it + GSTAGES < nk) load_stage(stage, (it + GSTAGES) * KC);
        cp_async_commit();
    }
    const bool neg = (t.flags & GF_NEG) != 0, accum = (t.flags & GF_ACCUM) != 0;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = wm * 32 + mb * 8 + g, col = wn * WN + nb * 8 + 2 * tq + e;
                if (row < t.m && col < t.n) {
                    double* c = t.C + row + (size_t)col * t.ldc;
                    double v = neg ? -acc[mb][nb][e] : acc[mb][nb][e];
                    if (accum) v += *c;
                    *c = v;
                }
            }
}

// LDL^T of the nb x nb diagonal block (held in shared memory, redundantly per CTA) and
// L_ik = F_ik L_kk^-T D^-1 for the task's rows.  The factored diagonal block goes to the solve
// panel S (its diagonal blocks are free until the inversion phase), never back to L, so CTAs of
// the same block column can read the un-factored block concurrently.
__global__ void __launch_bounds__(PANEL_ROWS) k_panel(const PanelTask* __restrict__ tasks, double* __restrict__ Lbase,
                                                      double* __restrict__ Sbase, unsigned long long* __restrict__ stats) {
    const PanelTask t = tasks[blockIdx.x];
    __shared__ double Dg[NB][NB + 1];
    __shared__ double dd[NB];
    const int tid = threadIdx.x;
    double* L = t.L;
    const size_t ld = (size_t)t.ld;
    for (int idx = tid; idx < NB * NB; idx += PANEL_ROWS) {
        const int r = idx % NB, c = idx / NB;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < t.nb && c < t.nb && r >= c) v = L[(t.k0 + r) + (size_t)(t.k0 + c) * ld];
        Dg[r][c] = v;
    }
    __syncthreads();
    if (tid < 32) {
        // warp 0: lane r holds row r of the (symmetric) block in registers; column c is eliminated
        // with shuffles: A[r][c2] -= (A[r][c] / A[c][c]) * A[c][c2]  for r, c2 > c.
        const int lane = tid;
        double row[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) row[c] = (lane >= c) ? Dg[lane][c] : Dg[c][lane];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            const double dc = __shfl_sync(0xffffffffu, row[c], c);
            const bool act = lane > c;
            const double lr = act ? row[c] / dc : 0.0;
#pragma unroll
            for (int c2 = c + 1; c2 < NB; ++c2) {
                const double pc = __shfl_sync(0xffffffffu, row[c2], c);
                row[c2] -= lr * pc;
            }
            if (act) Dg[lane][c] = lr;
            if (lane == c) dd[c] = dc;
        }
    }
    __syncthreads();
    if (tid < t.nr) {
        const size_t r = (size_t)t.r0 + tid;
        double f[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) f[c] = (c < t.nb) ? L[r + (size_t)(t.k0 + c) * ld] : 0.0;
#pragma unroll
        for (int c = 1; c < NB; ++c) {
            double yv = f[c];
#pragma unroll
            for (int c2 = 0; c2 < c; ++c2) yv -= f[c2] * Dg[c][c2];
            f[c] = yv;
        }
#pragma unroll
        for (int c = 0; c < NB; ++c)
            if (c < t.nb) L[r + (size_t)(t.k0 + c) * ld] = f[c] / dd[c];
    }
    if (t.write_diag) {
        double* Sd = Sbase + (L - Lbase);
        for (int idx = tid; idx < NB * NB; idx += PANEL_ROWS) {
            const int r = idx % NB, c = idx / NB;
            if (r < t.nb && c < t.nb && r >= c)
                Sd[(t.k0 + r) + (size_t)(t.k0 + c) * ld] = (r == c) ? dd[c] : Dg[r][c];
        }
        if (tid < t.nb) {
            const double dv = dd[tid];
            t.d[tid] = dv;
            const double av = fabs(dv);
            if (dv < 0.0) atomicAdd(&stats[0], 1ull);
            if (!(av > 0.0) || !isfinite(dv)) atomicAdd(&stats[1], 1ull);
            else {
                atomicMin(&stats[2], (unsigned long long)__double_as_longlong(av));
                atomicMax(&stats[3], (unsigned long long)__double_as_longlong(av));
            }
        }
    }
}

// X = L_kk^-1 (unit lower), in place in the solve panel; the diagonal (which held d) becomes 1.
__global__ void __launch_bounds__(32) k_diag_inverse(const InvTask* __restrict__ tasks) {
    const InvTask t = tasks[blockIdx.x];
    __shared__ double Lk[NB][NB + 1];
    const int lane = threadIdx.x;
    for (int c = 0; c < NB; ++c) {
        double v = 0.0;
        if (lane < t.nb && c < t.nb && lane > c) v = t.S[lane + (size_t)c * t.ld];
        Lk[lane][c] = v;
    }
    __syncwarp();
    // lane j owns column j of X
    double xcol[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) xcol[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 1; i < NB; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) s += Lk[i][k] * xcol[k];
        if (i > lane) xcol[i] = -s;
    }
    if (lane < t.nb) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (i < t.nb) t.S[i + (size_t)lane * t.ld] = xcol[i];
    }
}

struct FrontView {
    const int32_t* p;
    const int32_t* q;
    const int64_t* start;
    const int64_t* Loff;
    const int64_t* Uoff;
    const int64_t* woff;
    const int64_t* bndoff;
    const int32_t* bnd;
    const int32_t* rel;
};

// destination of every CSR slot inside the front panels (entries (i,j) with new(i) >= new(j)
// belong to the panel of the owner of j); -1 for the mirrored upper entries.
__global__ void k_dest_map(int64_t N, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                           const int32_t* __restrict__ iperm, const int32_t* __restrict__ node_of_new, FrontView fv,
                           int64_t* __restrict__ dest) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int32_t ni = iperm[i];
    for (int s = rowptr[i]; s < rowptr[i + 1]; ++s) {
        const int32_t nj = iperm[colidx[s]];
        int64_t dst = -1;
        if (ni >= nj) {
            const int32_t t = node_of_new[nj];
            const int64_t st = fv.start[t];
            const int32_t p = fv.p[t], q = fv.q[t];
            const int64_t col = nj - st;
            int64_t r;
            if (ni < st + p) r = ni - st;
            else {
                const int32_t* b = fv.bnd + fv.bndoff[t];
                int lo = 0, hi = q;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (b[mid] < ni) lo = mid + 1;
                    else hi = mid;
                }
                r = (lo < q && b[lo] == ni) ? p + lo : -1;
            }
            if (r >= 0) dst = fv.Loff[t] + r + col * (int64_t)(p + q);
            else dst = -2;   // structural inconsistency (reported by the host)
        }
        dest[s] = dst;
    }
}

__global__ void k_scatter(int64_t nnz, const int64_t* __restrict__ dest, const double* __restrict__ A,
                          const double* __restrict__ B, double sigma, double* __restrict__ L) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; s < nnz; s += stride) {
        const int64_t dst = dest[s];
        if (dst >= 0) L[dst] = B ? A[s] - sigma * B[s] : A[s];
    }
}

// child update matrix (lower triangle of U_c) added into the parent's panel / update matrix.
__global__ void __launch_bounds__(256) k_extend_add(const ExtTask* __restrict__ tasks, FrontView fv,
                                                    double* __restrict__ Lb, double* __restrict__ Ub) {
    const ExtTask t = tasks[blockIdx.x];
    const int c = t.child, par = c >> 1;
    const int qc = fv.q[c];
    const int pp = fv.p[par], qp = fv.q[par], mp = pp + qp;
    const int32_t* rel = fv.rel + fv.bndoff[c];
    const double* Uc = Ub + fv.Uoff[c];
    double* Lp = Lb + fv.Loff[par];
    double* Up = Ub + fv.Uoff[par];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = t.ti * 32 + tx;
    if (i >= qc) return;
    const int ri = rel[i];
    for (int jj = ty; jj < 32; jj += 8) {
        const int j = t.tj * 32 + jj;
        if (j > i || j >= qc) continue;
        const int rj = rel[j];
        const double v = Uc[i + (size_t)j * qc];
        if (rj < pp) Lp[ri + (size_t)rj * mp] += v;
        else Up[(ri - pp) + (size_t)(rj - pp) * qp] += v;
    }
}

// ---- solve sweeps -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fwd_gather(int64_t first, int has_children, FrontView fv,
                                                    const double* __restrict__ bperm, double* __restrict__ Wb) {
    const int64_t t = first + blockIdx.x;
    const int p = fv.p[t], q = fv.q[t], m = p + q;
    double* w = Wb + fv.woff[t];
    const int64_t st = fv.start[t];
    for (int i = threadIdx.x; i < m; i += blockDim.x) w[i] = (i < p) ? bperm[st + i] : 0.0;
    if (!has_children) return;
    for (int cc = 0; cc < 2; ++cc) {
        __syncthreads();
        const int64_t c = 2 * t + cc;
        const int pc = fv.p[c], qc = fv.q[c];
        const double* wc = Wb + fv.woff[c] + pc;
        const int32_t* rel = fv.rel + fv.bndoff[c];
        for (int j = threadIdx.x; j < qc; j += blockDim.x) w[rel[j]] += wc[j];
    }
}

// rows [r0, r0+nr) of  out = S * w_own ;  own rows -> y, boundary rows: w_bnd -= out.
// 32 rows x 32 k-lanes per CTA: every thread keeps 4 independent loads in flight.
// FWD_KL = 8 (256 threads) when a level has thousands of tiles, 32 (1024 threads) for the few
// large fronts near the root, where a tile's k loop would otherwise be long and serial.
template <int FWD_KL>
__global__ void __launch_bounds__(GEMV_ROWS * FWD_KL) k_fwd_gemv(const GemvTask* __restrict__ tasks) {
    const GemvTask t = tasks[blockIdx.x];
    __shared__ double red[FWD_KL][GEMV_ROWS + 1];
    const int lr = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int r = t.r0 + lr;
    const bool valid = lr < t.nr;
    const int kend_tile = min(t.p, (t.r0 < t.p) ? t.r0 + t.nr : t.p);
    double acc = 0.0;
    if (valid) {
        const int kend = (r < t.p) ? min(kend_tile, r + 1) : t.p;
        const double* Sr = t.S + r;
        const size_t ld = (size_t)t.ld;
        int k = kl;
        for (; k + 3 * FWD_KL < kend; k += 4 * FWD_KL) {
            const double s0 = Sr[(size_t)k * ld], s1 = Sr[(size_t)(k + FWD_KL) * ld];
            const double s2 = Sr[(size_t)(k + 2 * FWD_KL) * ld], s3 = Sr[(size_t)(k + 3 * FWD_KL) * ld];
            acc += s0 * t.w[k] + s1 * t.w[k + FWD_KL] + s2 * t.w[k + 2 * FWD_KL] + s3 * t.w[k + 3 * FWD_KL];
        }
        for (; k < kend; k += FWD_KL) acc += Sr[(size_t)k * ld] * t.w[k];
    }
    red[kl][lr] = acc;
    __syncthreads();
    if (kl == 0 && valid) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < FWD_KL; ++j) s += red[j][lr];
        if (r < t.p) t.y[r] = s;
        else t.w[r] -= s;
    }
}

__global__ void __launch_bounds__(256) k_bwd_gather(int64_t first, FrontView fv, const double* __restrict__ y,
                                                    const double* __restrict__ dinv, const double* __restrict__ x,
                                                    double* __restrict__ Wb) {
    const int64_t t = first + blockIdx.x;
    const int p = fv.p[t], q = fv.q[t], m = p + q;
    double* g = Wb + fv.woff[t];
    const int64_t st = fv.start[t];
    const int32_t* b = fv.bnd + fv.bndoff[t];
    for (int i = threadIdx.x; i < m; i += blockDim.x) g[i] = (i < p) ? y[st + i] * dinv[st + i] : -x[b[i - p]];
}

// columns [c0, c0+nc): x_own[c] = sum_{r>=c} S[r,c] * g[r].  4 warps share one column (each a
// quarter of the rows, 4 independent accumulators per lane), 8 columns per CTA.
template <int BWD_WPC>
__global__ void __launch_bounds__(32 * GEMVT_COLS * BWD_WPC) k_bwd_gemvT(const GemvTTask* __restrict__ tasks) {
    const GemvTTask t = tasks[blockIdx.x];
    __shared__ double red[GEMVT_COLS][BWD_WPC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wc = warp / BWD_WPC, part = warp % BWD_WPC;
    double acc = 0.0;
    if (wc < t.nc) {
        const int c = t.c0 + wc;
        const double* Sc = t.S + (size_t)c * t.ld;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int step = 32 * BWD_WPC;
        int r = c + part * 32 + lane;
        for (; r + 3 * step < t.m; r += 4 * step) {
            a0 += Sc[r] * t.g[r];
            a1 += Sc[r + step] * t.g[r + step];
            a2 += Sc[r + 2 * step] * t.g[r + 2 * step];
            a3 += Sc[r + 3 * step] * t.g[r + 3 * step];
        }
        for (; r < t.m; r += step) a0 += Sc[r] * t.g[r];
        acc = (a0 + a1) + (a2 + a3);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (lane == 0 && wc < GEMVT_COLS) red[wc][part] = acc;
    __syncthreads();
    if (threadIdx.x < t.nc) {
        double sacc = 0.0;
#pragma unroll
        for (int j = 0; j < BWD_WPC; ++j) sacc += red[threadIdx.x][j];
        t.x[t.c0 + threadIdx.x] = sacc;
    }
}

// ---- small fronts (m <= SMALL_M): one CTA does a whole front, work vector in shared memory ----
constexpr int SMALL_M = 256;
__global__ void __launch_bounds__(128) k_fwd_small(int64_t first, int has_children, FrontView fv,
                                                   const double* __restrict__ bperm, double* __restrict__ Wb,
                                                   const double* __restrict__ Sb, double* __restrict__ y) {
    const int64_t t = first + blockIdx.x;
    const int p = fv.p[t], q = fv.q[t], m = p + q;
    __shared__ double sw[SMALL_M];
    double* w = Wb + fv.woff[t];
    const int64_t st = fv.start[t];
    for (int i = threadIdx.x; i < m; i += 128) sw[i] = (i < p) ? bperm[st + i] : 0.0;
    if (has_children) {
        for (int cc = 0; cc < 2; ++cc) {
            __syncthreads();
            const int64_t c = 2 * t + cc;
            const int pc = fv.p[c], qc = fv.q[c];
            const double* wc = Wb + fv.woff[c] + pc;
            const int32_t* rel = fv.rel + fv.bndoff[c];
            for (int j = threadIdx.x; j < qc; j += 128) sw[rel[j]] += wc[j];
        }
    }
    __syncthreads();
    const double* S = Sb + fv.Loff[t];
    for (int r = threadIdx.x; r < m; r += 128) {
        const int kend = (r < p) ? r + 1 : p;
        const double* Sr = S + r;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = 0;
        for (; k + 3 < kend; k += 4) {
            a0 += Sr[(size_t)k * m] * sw[k];
            a1 += Sr[(size_t)(k + 1) * m] * sw[k + 1];
            a2 += Sr[(size_t)(k + 2) * m] * sw[k + 2];
            a3 += Sr[(size_t)(k + 3) * m] * sw[k + 3];
        }
        for (; k < kend; ++k) a0 += Sr[(size_t)k * m] * sw[k];
        const double acc = (a0 + a1) + (a2 + a3);
        if (r < p) y[st + r] = acc;
        else w[r] = sw[r] - acc;
    }
}

__global__ void __launch_bounds__(128) k_bwd_small(int64_t first, FrontView fv, const double* __restrict__ y,
                                                   const double* __restrict__ dinv, double* __restrict__ x,
                                                   const double* __restrict__ Sb) {
    const int64_t t = first + blockIdx.x;
    const int p = fv.p[t], q = fv.q[t], m = p + q;
    __shared__ double sg[SMALL_M];
    const int64_t st = fv.start[t];
    const int32_t* b = fv.bnd + fv.bndoff[t];
    for (int i = threadIdx.x; i < m; i += 128) sg[i] = (i < p) ? y[st + i] * dinv[st + i] : -x[b[i - p]];
    __syncthreads();
    const double* S = Sb + fv.Loff[t];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = warp; c < p; c += 4) {
        const double* Sc = S + (size_t)c * m;
        double a0 = 0.0, a1 = 0.0;
        int r = c + lane;
        for (; r + 32 < m; r += 64) {
            a0 += Sc[r] * sg[r];
            a1 += Sc[r + 32] * sg[r + 32];
        }
        if (r < m) a0 += Sc[r] * sg[r];
        double acc = a0 + a1;
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) x[st + c] = acc;
    }
}

__global__ void k_permute_in(int64_t N, const int32_t* __restrict__ perm, const double* __restrict__ b, double* __restrict__ bp) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) bp[i] = b[perm[i]];
}
__global__ void k_permute_out(int64_t N, const int32_t* __restrict__ perm, const double* __restrict__ x, double* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) out[perm[i]] = x[i];
}
__global__ void k_recip(int64_t N, const double* __restrict__ d, double* __restrict__ dinv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) dinv[i] = 1.0 / d[i];
}
__global__ void k_count_bad_dest(int64_t nnz, const int64_t* __restrict__ dest, unsigned long long* __restrict__ out) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s < nnz && dest[s] == -2) atomicAdd(out, 1ull);
}

// ---------------------------------------------------------------------------------------------
// host: task generation
// ---------------------------------------------------------------------------------------------
template <class T>
static void upload(DevBuf<T>& buf, const std::vector<T>& v, cudaStream_t st) {
    buf.alloc(v.size(), st);
    if (!v.empty()) WS_CUDA(cudaMemcpyAsync(buf.get(), v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

static FrontView view(const ws_solver* s) {
    return FrontView{s->fr_p.get(), s->fr_q.get(), s->fr_start.get(), s->fr_Loff.get(), s->fr_Uoff.get(),
                     s->fr_woff.get(), s->fr_bndoff.get(), s->bnd.get(), s->rel.get()};
}

static void build_plan(ws_solver* s, cudaStream_t st) {
    const Symbolic& Y = s->sym;
    const int depth = Y.depth;
    std::vector<GemmTask> G;
    std::vector<PanelTask> P;
    std::vector<ExtTask> E;
    std::vector<InvTask> I;
    std::vector<GemvTask> F;
    std::vector<GemvTTask> Bk;
    double* Lb = s->L.get();
    double* Sb = s->S.get();
    double* Tb = s->T.get();
    double* Ub = s->U.get();
    double* Wb = s->W.get();
    double* db = s->d.get();
    s->levels.assign(depth + 1, LevelPlan());
    for (int l = depth; l >= 0; --l) {
        LevelPlan& LP = s->levels[l];
        const int64_t f0 = (int64_t)1 << l, f1 = (int64_t)1 << (l + 1);
        LP.first_front = f0;
        LP.nfronts = f1 - f0;
        int maxp = 0;
        for (int64_t t = f0; t < f1; ++t) maxp = std::max(maxp, (int)Y.p[t]);
        const int nkb = (maxp + NB - 1) / NB;
        LP.upd.assign(nkb, Range());
        LP.pan.assign(nkb, Range());
        LP.trail.assign(nkb, Range());
        int maxm = 0;
        for (int64_t t = f0; t < f1; ++t) maxm = std::max(maxm, (int)(Y.p[t] + Y.q[t]));
        LP.small = maxm <= SMALL_M;
        // few large fronts: a left-looking block-column update has too few tiles to fill the GPU
        // (and a long serial k loop); update the whole trailing panel after every block column.
        LP.right_looking = (f1 - f0) <= 64 && maxp > 4 * NB;
        // extend-add of the children of this level's fronts (left children first, then right)
        if (l < depth) {
            for (int side = 0; side < 2; ++side) {
                Range& R = side ? LP.ext_right : LP.ext_left;
                R.off = (int64_t)E.size();
                for (int64_t t = f0; t < f1; ++t) {
                    const int64_t c = 2 * t + side;
                    const int nt = (Y.q[c] + 31) / 32;
                    for (int ti = 0; ti < nt; ++ti)
                        for (int tj = 0; tj <= ti; ++tj) E.push_back(ExtTask{(int32_t)c, ti, tj});
                }
                R.cnt = (int64_t)E.size() - R.off;
            }
        }
        for (int kb = 0; kb < nkb; ++kb) {
            const int k0 = kb * NB;
            LP.upd[kb].off = (int64_t)G.size();
            if (kb > 0 && !LP.right_looking)
                for (int64_t t = f0; t < f1; ++t) {
                    const int p = Y.p[t], m = p + Y.q[t];
                    if (p <= k0) continue;
                    const int nbk = std::min(NB, p - k0);
                    double* Lt = Lb + Y.Loff[t];
                    for (int r0 = k0; r0 < m; r0 += GT) {
                        GemmTask g{};
                        g.A = Lt + r0;
                        g.B = Lt + k0;
                        g.C = Lt + r0 + (size_t)k0 * m;
                        g.dscale = db + Y.start[t];
                        g.lda = g.ldb = g.ldc = m;
                        g.m = std::min(GT, m - r0);
                        g.n = nbk;
                        g.k = k0;
                        g.flags = GF_ACCUM | GF_NEG;
                        G.push_back(g);
                    }
                }
            LP.upd[kb].cnt = (int64_t)G.size() - LP.upd[kb].off;
            LP.pan[kb].off = (int64_t)P.size();
            for (int64_t t = f0; t < f1; ++t) {
                const int p = Y.p[t], m = p + Y.q[t];
                if (p <= k0) continue;
                const int nbk = std::min(NB, p - k0);
                double* Lt = Lb + Y.Loff[t];
                bool first = true;
                int r0 = k0 + nbk;
                do {
                    PanelTask pt{};
                    pt.L = Lt;
                    pt.d = db + Y.start[t] + k0;
                    pt.ld = m;
                    pt.k0 = k0;
                    pt.nb = nbk;
                    pt.r0 = r0;
                    pt.nr = std::max(0, std::min(PANEL_ROWS, m - r0));
                    pt.write_diag = first ? 1 : 0;
                    P.push_back(pt);
                    first = false;
                    r0 += PANEL_ROWS;
                } while (r0 < m);
            }
            LP.pan[kb].cnt = (int64_t)P.size() - LP.pan[kb].off;
            LP.trail[kb].off = (int64_t)G.size();
            if (LP.right_looking)
                for (int64_t t = f0; t < f1; ++t) {
                    const int p = Y.p[t], m = p + Y.q[t];
                    if (p <= k0) continue;
                    const int nbk = std::min(NB, p - k0);
                    const int c0 = k0 + nbk;           // first trailing row / column
                    double* Lt = Lb + Y.Loff[t];
                    for (int i0 = c0; i0 < m; i0 += GT)
                        for (int j0 = c0; j0 <= i0 && j0 < p; j0 += GT) {
                            GemmTask g{};
                            g.A = Lt + i0 + (size_t)k0 * m;
                            g.B = Lt + j0 + (size_t)k0 * m;
                            g.C = Lt + i0 + (size_t)j0 * m;
                            g.dscale = db + Y.start[t] + k0;
                            g.lda = g.ldb = g.ldc = m;
                            g.m = std::min(GT, m - i0);
                            g.n = std::min(GT, p - j0);
                            g.k = nbk;
                            g.flags = GF_ACCUM | GF_NEG;
                            G.push_back(g);
                        }
                }
            LP.trail[kb].cnt = (int64_t)G.size() - LP.trail[kb].off;
        }
        // Schur complement U -= L21 D L21^T (lower tiles)
        LP.syrk.off = (int64_t)G.size();
        for (int64_t t = f0; t < f1; ++t) {
            const int p = Y.p[t], q = Y.q[t], m = p + q;
            if (!p || !q) continue;
            double* Lt = Lb + Y.Loff[t];
            double* Ut = Ub + Y.Uoff[t];
            for (int i0 = 0; i0 < q; i0 += GT)
                for (int j0 = 0; j0 <= i0; j0 += GT) {
                    GemmTask g{};
                    g.A = Lt + p + i0;
                    g.B = Lt + p + j0;
                    g.C = Ut + i0 + (size_t)j0 * q;
                    g.dscale = db + Y.start[t];
                    g.lda = g.ldb = m;
                    g.ldc = q;
                    g.m = std::min(GT, q - i0);
                    g.n = std::min(GT, q - j0);
                    g.k = p;
                    g.flags = GF_ACCUM | GF_NEG;
                    G.push_back(g);
                }
        }
        LP.syrk.cnt = (int64_t)G.size() - LP.syrk.off;
        // solve sweeps
        LP.fwd.off = (int64_t)F.size();
        LP.bwd.off = (int64_t)Bk.size();
        for (int64_t t = f0; t < f1 && !LP.small; ++t) {
            const int p = Y.p[t], q = Y.q[t], m = p + q;
            if (!p) continue;
            for (int r0 = 0; r0 < m; r0 += GEMV_ROWS) {
                // a tile must not straddle the own / boundary split (different k extents)
                int nr = std::min(GEMV_ROWS, m - r0);
                if (r0 < p && r0 + nr > p) nr = p - r0;
                GemvTask f{};
                f.S = Sb + Y.Loff[t];
                f.w = Wb + Y.woff[t];
                f.y = s->y.get() + Y.start[t];
                f.ld = m; f.p = p; f.r0 = r0; f.nr = nr;
                F.push_back(f);
                if (nr < GEMV_ROWS && r0 + nr < m) {     // remainder of a straddling tile
                    GemvTask f2 = f;
                    f2.r0 = r0 + nr;
                    f2.nr = std::min(GEMV_ROWS - nr, m - f2.r0);
                    F.push_back(f2);
                }
            }
            for (int c0 = 0; c0 < p; c0 += GEMVT_COLS) {
                GemvTTask b{};
                b.S = Sb + Y.Loff[t];
                b.g = Wb + Y.woff[t];
                b.x = s->x.get() + Y.start[t];
                b.ld = m; b.m = m; b.c0 = c0; b.nc = std::min(GEMVT_COLS, p - c0);
                Bk.push_back(b);
            }
        }
        LP.fwd.cnt = (int64_t)F.size() - LP.fwd.off;
        LP.bwd.cnt = (int64_t)Bk.size() - LP.bwd.off;
    }
    // ---- selective inversion (all fronts, independent of the tree)
    s->inv0.off = 0;
    for (int64_t t = 1; t < Y.nfronts; ++t) {
        const int p = Y.p[t], m = p + Y.q[t];
        for (int k0 = 0; k0 < p; k0 += NB)
            I.push_back(InvTask{Sb + Y.Loff[t] + k0 + (size_t)k0 * m, m, std::min(NB, p - k0)});
    }
    s->inv0.cnt = (int64_t)I.size();
    s->inv_stages.clear();
    for (int h = NB; h < Y.maxp; h *= 2) {
        Range r1, r2;
        r1.off = (int64_t)G.size();
        for (int64_t t = 1; t < Y.nfronts; ++t) {
            const int p = Y.p[t], m = p + Y.q[t];
            const double* Lt = Lb + Y.Loff[t];
            double* St = Sb + Y.Loff[t];
            double* Tt = Tb + Y.Loff[t];
            for (int r0 = 0; r0 + h < p; r0 += 2 * h) {
                const int hB = std::min(h, p - r0 - h);
                // Tmp(hB x h) = L_BA * X_AA ; X_AA lower triangular -> k starts at the tile column
                for (int i0 = 0; i0 < hB; i0 += GT)
                    for (int j0 = 0; j0 < h; j0 += GT) {
                        GemmTask g{};
                        g.A = Lt + (r0 + h + i0) + (size_t)(r0 + j0) * m;
                        g.B = St + (r0 + j0) + (size_t)(r0 + j0) * m;
                        g.C = Tt + (r0 + h + i0) + (size_t)(r0 + j0) * m;
                        g.lda = g.ldb = g.ldc = m;
                        g.m = std::min(GT, hB - i0);
                        g.n = std::min(GT, h - j0);
                        g.k = h - j0;
                        g.flags = GF_BN;
                        G.push_back(g);
                    }
            }
        }
        r1.cnt = (int64_t)G.size() - r1.off;
        r2.off = (int64_t)G.size();
        for (int64_t t = 1; t < Y.nfronts; ++t) {
            const int p = Y.p[t], m = p + Y.q[t];
            double* St = Sb + Y.Loff[t];
            const double* Tt = Tb + Y.Loff[t];
            for (int r0 = 0; r0 + h < p; r0 += 2 * h) {
                const int hB = std::min(h, p - r0 - h);
                // X_BA = - X_BB * Tmp ; X_BB lower triangular -> k ends with the tile row
                for (int i0 = 0; i0 < hB; i0 += GT)
                    for (int j0 = 0; j0 < h; j0 += GT) {
                        GemmTask g{};
                        g.A = St + (r0 + h + i0) + (size_t)(r0 + h) * m;
                        g.B = Tt + (r0 + h) + (size_t)(r0 + j0) * m;
                        g.C = St + (r0 + h + i0) + (size_t)(r0 + j0) * m;
                        g.lda = g.ldb = g.ldc = m;
                        g.m = std::min(GT, hB - i0);
                        g.n = std::min(GT, h - j0);
                        g.k = std::min(hB, i0 + GT);
                        g.flags = GF_BN | GF_NEG;
                        G.push_back(g);
                    }
            }
        }
        r2.cnt = (int64_t)G.size() - r2.off;
        s->inv_stages.push_back({r1, r2});
    }
    // W21 = L21 * X
    s->w21.off = (int64_t)G.size();
    for (int64_t t = 1; t < Y.nfronts; ++t) {
        const int p = Y.p[t], q = Y.q[t], m = p + q;
        if (!p || !q) continue;
        const double* Lt = Lb + Y.Loff[t];
        double* St = Sb + Y.Loff[t];
        for (int i0 = 0; i0 < q; i0 += GT)
            for (int j0 = 0; j0 < p; j0 += GT) {
                GemmTask g{};
                g.A = Lt + (p + i0) + (size_t)j0 * m;
                g.B = St + j0 + (size_t)j0 * m;
                g.C = St + (p + i0) + (size_t)j0 * m;
                g.lda = g.ldb = g.ldc = m;
                g.m = std::min(GT, q - i0);
                g.n = std::min(GT, p - j0);
                g.k = p - j0;
                g.flags = GF_BN;
                G.push_back(g);
            }
    }
    s->w21.cnt = (int64_t)G.size() - s->w21.off;
    upload(s->gemm_tasks, G, st);
    upload(s->panel_tasks, P, st);
    upload(s->ext_tasks, E, st);
    upload(s->inv_tasks, I, st);
    upload(s->gemv_tasks, F, st);
    upload(s->gemvt_tasks, Bk, st);
    WS_CUDA(cudaStreamSynchronize(st));   // the host vectors die here
    s->bytes += (int64_t)(G.size() * sizeof(GemmTask) + P.size() * sizeof(PanelTask) + E.size() * sizeof(ExtTask) +
                          I.size() * sizeof(InvTask) + F.size() * sizeof(GemvTask) + Bk.size() * sizeof(GemvTTask));
}

static void launch_gemm(const ws_solver* s, const Range& r, bool narrow, cudaStream_t st) {
    if (!r.cnt) return;
    static bool configured = false;
    if (!configured) {
        WS_CUDA(cudaFuncSetAttribute(k_gemm_tiles<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<32>()));
        WS_CUDA(cudaFuncSetAttribute(k_gemm_tiles<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()));
        configured = true;
    }
    if (narrow) k_gemm_tiles<32><<<(unsigned)r.cnt, 128, gemm_smem_bytes<32>(), st>>>(s->gemm_tasks.get() + r.off);
    else k_gemm_tiles<64><<<(unsigned)r.cnt, 128, gemm_smem_bytes<64>(), st>>>(s->gemm_tasks.get() + r.off);
    count_launch(1);
}

}  // namespace ws

using namespace ws;

extern "C" {

int ws_solver_create(const ws_pattern* p, const ws_symbolic* sym, ws_solver** out, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(out, "out == NULL");
    *out = nullptr;
    WS_REQUIRE(p && sym, "null handle");
    WS_REQUIRE(p->N == sym->S.N, "pattern and symbolic analysis have different sizes");
    DeviceGuard g(device);
    configure_mempool(device);
    cudaStream_t st = (cudaStream_t)stream;
    auto* s = new ws_solver();
    const bool trace = getenv("WS_TRACE") != nullptr;
    auto tnow = [&]() {
        if (trace) cudaStreamSynchronize(st);
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    const double tc0 = tnow();
    try {
        const Symbolic& Y = sym->S;
        s->device = device;
        s->sym = Y;
        s->N = Y.N;
        s->nnz = p->nnz;
        s->nfronts = Y.nfronts;
        s->depth = Y.depth;
        upload(s->perm, Y.perm, st);
        upload(s->iperm, Y.iperm, st);
        upload(s->node_of_new, Y.node_of_new, st);
        upload(s->fr_p, Y.p, st);
        upload(s->fr_q, Y.q, st);
        upload(s->fr_start, Y.start, st);
        upload(s->fr_Loff, Y.Loff, st);
        upload(s->fr_Uoff, Y.Uoff, st);
        upload(s->fr_woff, Y.woff, st);
        upload(s->fr_bndoff, Y.bndoff, st);
        upload(s->bnd, Y.bnd, st);
        upload(s->rel, Y.rel, st);
        s->L.alloc(Y.Ltotal, st);
        s->S.alloc(Y.Ltotal, st);
        s->T.alloc(Y.Ltotal, st);
        s->U.alloc(std::max<int64_t>(Y.Utotal, 2), st);
        s->W.alloc(Y.Wtotal, st);
        s->d.alloc(Y.N, st);
        s->dinv.alloc(Y.N, st);
        s->y.alloc(Y.N, st);
        s->x.alloc(Y.N, st);
        s->bperm.alloc(Y.N, st);
        s->stats.alloc(4, st);
        s->dest.alloc(p->nnz, st);
        s->bytes = (3 * Y.Ltotal + Y.Utotal + Y.Wtotal + 5 * Y.N) * 8 + p->nnz * 8;
        const double tc1 = tnow();
        const int T = 128;
        k_dest_map<<<ceil_div(Y.N, T), T, 0, st>>>(Y.N, p->rowptr.get(), p->colidx.get(), s->iperm.get(),
                                                     s->node_of_new.get(), view(s), s->dest.get());
        WS_CUDA(cudaMemsetAsync(s->stats.get(), 0, 4 * sizeof(unsigned long long), st));
        k_count_bad_dest<<<ceil_div(p->nnz, 256), 256, 0, st>>>(p->nnz, s->dest.get(), s->stats.get());
        unsigned long long bad = 0;
        WS_CUDA(cudaMemcpyAsync(&bad, s->stats.get(), sizeof bad, cudaMemcpyDeviceToHost, st));
        count_launch(2);
        const double tc2 = tnow();
        build_plan(s, st);   // synchronises the stream
        if (trace)
            fprintf(stderr, "[ws_solver_create] copy+upload+alloc %.1f ms, dest map %.1f ms, plan %.1f ms\n",
                    (tc1 - tc0) * 1e3, (tc2 - tc1) * 1e3, (tnow() - tc2) * 1e3);
        WS_CUDA(cudaGetLastError());
        if (bad) throw Error(WS_ERR_STATE, "pattern and symbolic analysis are inconsistent (matrix entry outside its front)");
    } catch (...) {
        delete s;
        throw;
    }
    *out = s;
    WS_API_END
}

int ws_solver_factor(ws_solver* s, const double* A_vals, const double* B_vals, double sigma, int device,
                     void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(s && A_vals, "null pointer");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const Symbolic& Y = s->sym;
    s->factored = false;
    WS_CUDA(cudaMemsetAsync(s->L.get(), 0, s->L.bytes(), st));
    WS_CUDA(cudaMemsetAsync(s->U.get(), 0, s->U.bytes(), st));
    WS_CUDA(cudaMemsetAsync(s->S.get(), 0, s->S.bytes(), st));   // blocks above the block diagonal are read as zeros
    unsigned long long init[4] = {0ull, 0ull, ~0ull, 0ull};
    WS_CUDA(cudaMemcpyAsync(s->stats.get(), init, sizeof init, cudaMemcpyHostToDevice, st));
    {
        int blocks = (int)std::min<int64_t>((s->nnz + 255) / 256, (int64_t)kNumSMs * 32);
        k_scatter<<<blocks, 256, 0, st>>>(s->nnz, s->dest.get(), A_vals, B_vals, sigma, s->L.get());
        count_launch(1);
    }
    FrontView fv = view(s);
    for (int l = Y.depth; l >= 0; --l) {
        const LevelPlan& LP = s->levels[l];
        if (LP.ext_left.cnt) {
            k_extend_add<<<(unsigned)LP.ext_left.cnt, 256, 0, st>>>(s->ext_tasks.get() + LP.ext_left.off, fv, s->L.get(), s->U.get());
            count_launch(1);
        }
        if (LP.ext_right.cnt) {
            k_extend_add<<<(unsigned)LP.ext_right.cnt, 256, 0, st>>>(s->ext_tasks.get() + LP.ext_right.off, fv, s->L.get(), s->U.get());
            count_launch(1);
        }
        for (size_t kb = 0; kb < LP.pan.size(); ++kb) {
            launch_gemm(s, LP.upd[kb], true, st);
            if (LP.pan[kb].cnt) {
                k_panel<<<(unsigned)LP.pan[kb].cnt, PANEL_ROWS, 0, st>>>(s->panel_tasks.get() + LP.pan[kb].off, s->L.get(),
                                                                         s->S.get(), s->stats.get());
                count_launch(1);
            }
            launch_gemm(s, LP.trail[kb], false, st);
        }
        launch_gemm(s, LP.syrk, false, st);
    }
    // selective inversion: S_top = L11^-1, S_bottom = L21 * L11^-1
    if (s->inv0.cnt) {
        k_diag_inverse<<<(unsigned)s->inv0.cnt, 32, 0, st>>>(s->inv_tasks.get());
        count_launch(1);
    }
    for (auto& stg : s->inv_stages) {
        launch_gemm(s, stg.first, false, st);
        launch_gemm(s, stg.second, false, st);
    }
    launch_gemm(s, s->w21, false, st);
    k_recip<<<ceil_div(s->N, 256), 256, 0, st>>>(s->N, s->d.get(), s->dinv.get());
    count_launch(1);
    unsigned long long h[4];
    WS_CUDA(cudaMemcpyAsync(h, s->stats.get(), sizeof h, cudaMemcpyDeviceToHost, st));
    WS_CUDA(cudaStreamSynchronize(st));
    WS_CUDA(cudaGetLastError());
    s->h_neg = (int64_t)h[0];
    s->h_bad = (int64_t)h[1];
    long long mn = (long long)h[2], mx = (long long)h[3];
    std::memcpy(&s->h_minpiv, &mn, 8);
    std::memcpy(&s->h_maxpiv, &mx, 8);
    if (h[2] == ~0ull) s->h_minpiv = 0.0;
    if (s->h_bad)
        throw Error(WS_ERR_NUMERIC, "factorisation met " + std::to_string(s->h_bad) + " zero / non-finite pivots (matrix singular at this shift)");
    s->factored = true;
    WS_API_END
}

int ws_solver_solve(ws_solver* s, const double* b, double* x, int nrhs, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(s && b && x && nrhs >= 1, "bad arguments");
    if (!s->factored) throw Error(WS_ERR_STATE, "ws_solver_solve before a successful ws_solver_factor");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const Symbolic& Y = s->sym;
    FrontView fv = view(s);
    const int T = 256;
    for (int rhs = 0; rhs < nrhs; ++rhs) {
        const double* bi = b + (size_t)rhs * s->N;
        double* xi = x + (size_t)rhs * s->N;
        k_permute_in<<<ceil_div(s->N, T), T, 0, st>>>(s->N, s->perm.get(), bi, s->bperm.get());
        int nl = 2;
        for (int l = Y.depth; l >= 0; --l) {
            const LevelPlan& LP = s->levels[l];
            const int hc = l < Y.depth ? 1 : 0;
            if (LP.small) {
                k_fwd_small<<<(unsigned)LP.nfronts, 128, 0, st>>>(LP.first_front, hc, fv, s->bperm.get(), s->W.get(), s->S.get(), s->y.get());
                ++nl;
            } else {
                k_fwd_gather<<<(unsigned)LP.nfronts, 256, 0, st>>>(LP.first_front, hc, fv, s->bperm.get(), s->W.get());
                if (LP.fwd.cnt) {
                    if (LP.fwd.cnt >= 500) k_fwd_gemv<8><<<(unsigned)LP.fwd.cnt, GEMV_ROWS * 8, 0, st>>>(s->gemv_tasks.get() + LP.fwd.off);
                    else k_fwd_gemv<32><<<(unsigned)LP.fwd.cnt, GEMV_ROWS * 32, 0, st>>>(s->gemv_tasks.get() + LP.fwd.off);
                }
                nl += 2;
            }
        }
        for (int l = 0; l <= Y.depth; ++l) {
            const LevelPlan& LP = s->levels[l];
            if (LP.small) {
                k_bwd_small<<<(unsigned)LP.nfronts, 128, 0, st>>>(LP.first_front, fv, s->y.get(), s->dinv.get(), s->x.get(), s->S.get());
                ++nl;
            } else {
                k_bwd_gather<<<(unsigned)LP.nfronts, 256, 0, st>>>(LP.first_front, fv, s->y.get(), s->dinv.get(), s->x.get(), s->W.get());
                if (LP.bwd.cnt) {
                    if (LP.bwd.cnt >= 430) k_bwd_gemvT<1><<<(unsigned)LP.bwd.cnt, 32 * GEMVT_COLS, 0, st>>>(s->gemvt_tasks.get() + LP.bwd.off);
                    else k_bwd_gemvT<4><<<(unsigned)LP.bwd.cnt, 32 * GEMVT_COLS * 4, 0, st>>>(s->gemvt_tasks.get() + LP.bwd.off);
                }
                nl += 2;
            }
        }
        k_permute_out<<<ceil_div(s->N, T), T, 0, st>>>(s->N, s->perm.get(), s->x.get(), xi);
        count_launch(nl);
    }
    WS_CUDA(cudaGetLastError());
    WS_API_END
}

int ws_solver_stats(const ws_solver* s, int64_t* h_info, double* h_dinfo) {
    WS_API_BEGIN
    WS_REQUIRE(s, "solver == NULL");
    if (h_info) {
        h_info[0] = s->sym.nnzL; h_info[1] = s->nfronts - 1; h_info[2] = s->sym.maxfront; h_info[3] = s->bytes;
        h_info[4] = s->h_neg; h_info[5] = s->N - s->h_neg - s->h_bad; h_info[6] = s->h_bad; h_info[7] = s->sym.depth;
    }
    if (h_dinfo) {
        h_dinfo[0] = s->sym.flops; h_dinfo[1] = s->h_minpiv; h_dinfo[2] = s->h_maxpiv;
    }
    WS_API_END
}

int ws_solver_destroy(ws_solver* s) {
    WS_API_BEGIN
    if (s) {
        DeviceGuard g(s->device);
        delete s;
    }
    WS_API_END
}

}  // extern "C"
