// K8, numeric phase: multifrontal LDL^T on the device + selective inversion + solve sweeps.
//
// Everything is expressed as batched *tile tasks*: the host (ws_solver_create) walks the
// front tree once and emits task records for generic kernels --
//     k_gemm_tiles   C(64xBN tile) (+)= (-) A * op(B)        fp64 tensor cores (DMMA m8n8k4)
//     k_panel        NBxNB diagonal block LDL^T + row-block TRSM
//     k_diag_inverse unit-lower NBxNB inverse
//     k_extend_add   child update matrix -> parent front
//     k_fwd_front / k_fwd_tile / k_bwd_front / k_bwd_tile   the two solve sweeps as dense mat-vecs
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

struct FrontRec;
struct FwdTile;
struct BwdTile;

struct LevelPlan {
    std::vector<Range> upd, pan, trail;   // per block column (upd: left-looking, trail: right-looking)
    Range syrk, ext_left, ext_right, fwd, bwd;
    int64_t first_front = 0, nfronts = 0;
    bool small = false;                   // the level goes through the one-CTA-per-front solve kernels
    bool narrow_tiles = false;            // tile kernels in their 256-thread variants (a level of few fronts that would fit one CTA each)
    int nt = 32;                          // threads of those kernels: largest front rounded up to a warp
    int maxp = 0, maxm = 0;
    bool right_looking = false;
    bool fused = false;                   // factorised by k_front_fused (one CTA per front, no tile tasks)
    int64_t fwd_mid = 0, bwd_mid = 0;     // tile tasks of the left half of the level's fronts (sweep pipelining)
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
    ws::DevBuf<double> L, S, T, U, W, d, dinv, yd, x;
    ws::DevBuf<ws::FrontRec> recs;
    ws::DevBuf<int32_t> wsrc0, wsrc1;       // gather maps of the forward sweep
    ws::DevBuf<int32_t> sflags;             // per-front completion counters of the two sweeps [2 * nfronts]
    // [0] negative pivots [1] NaN/Inf pivots [2] min |pivot| bits [3] max |pivot| bits [4] perturbed pivots
    // [5] max |M_ij| bits (k_scatter: the scale of the static-pivot threshold) [6..8] probe norms |r|, |x|, |b| (bits)
    ws::DevBuf<unsigned long long> stats;
    // tasks
    ws::DevBuf<ws::GemmTask> gemm_tasks;
    ws::DevBuf<ws::PanelTask> panel_tasks;
    ws::DevBuf<ws::ExtTask> ext_tasks;
    ws::DevBuf<ws::InvTask> inv_tasks;
    ws::DevBuf<ws::FwdTile> fwd_tiles;
    ws::DevBuf<ws::BwdTile> bwd_tiles;
    std::vector<ws::LevelPlan> levels;      // index = tree level
    bool plan_built = false;                // tile tasks are emitted on the first factorisation (see build_plan)
    bool overlap = true;                    // sweeps chained with programmatic dependent launch + per-front flags (ws_solver_set_overlap)
    std::vector<ws::GemmTask> h_gemm;       // host copies stay alive: the uploads are asynchronous
    std::vector<ws::PanelTask> h_panel;
    std::vector<ws::ExtTask> h_ext;
    std::vector<ws::InvTask> h_inv;
    std::vector<ws::FwdTile> h_fwd;
    std::vector<ws::BwdTile> h_bwd;
    ws::Range inv0;
    std::vector<std::pair<ws::Range, ws::Range>> inv_stages;   // (step1, step2) per doubling stage
    ws::Range w21;
    bool factored = false;
    int64_t h_neg = 0, h_bad = 0;
    double h_minpiv = 0, h_maxpiv = 0;
    int64_t bytes = 0;
    // accuracy guard (static pivoting + probe + iterative refinement, see ws_solver_factor)
    const ws_pattern* pat = nullptr;        // the pattern M lives on (outlives the solver)
    const double* fA = nullptr;             // value arrays of the last factorisation (caller keeps them alive
    const double* fB = nullptr;             // while refined solves run: M x = fA x - fsigma fB x)
    double fsigma = 0.0;
    int refine_steps = 0;                   // refinement steps every solve performs (0: the factorisation passed the probe as is)
    int64_t h_pert = 0;                     // pivots replaced by +-tau
    int probe_iters = 0;
    double h_eta0 = 0.0, h_eta = 0.0, h_maxabs = 0.0;   // probe backward error before / after refinement; max |M_ij|
    ws::DevBuf<double> rf_b, rf_r, rf_y, rf_dx, rf_x;
    int64_t uid = 0;                        // unique per created solver (a work space that cached graphs for a solver must not
                                            // mistake a new solver at a recycled address for it)
    unsigned long long* h_stats = nullptr;  // pinned landing buffer of the statistics (factor_begin -> factor_end)
    bool factor_pending = false;
    ~ws_solver() {
        if (h_stats) cudaFreeHost(h_stats);
    }
};

namespace ws {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
constexpr int NSTATS = 12;
// static pivoting: |pivot| < PIV_REL * max|M_ij| is replaced by +-that (a diagonal backward error of 1e-10 relative;
// growth bounded by 1e10, which two refinement steps absorb); counted in stats[4]
constexpr double PIV_REL = 1e-10;
__device__ __forceinline__ double pivot_tau(const unsigned long long* stats) {
    return PIV_REL * __longlong_as_double((long long)stats[5]);
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
    // Epilogue in two passes: first ALL loads of the output tile, then the sums and the stores.  Written as one
    // read-modify-write per entry the compiler must keep every load behind the previous store (they might alias), and a thread's
    // 32 entries became 32 dependent global round trips -- 78 % of the samples of a small-grid launch sat on that line
    // (profiles/r02c_factor_levels.txt).
    double cold[4][NBLK][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = wm * 32 + mb * 8 + g, col = wn * WN + nb * 8 + 2 * tq + e;
                cold[mb][nb][e] = (accum && row < t.m && col < t.n) ? t.C[row + (size_t)col * t.ldc] : 0.0;
            }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = wm * 32 + mb * 8 + g, col = wn * WN + nb * 8 + 2 * tq + e;
                if (row < t.m && col < t.n) {
                    double v = neg ? -acc[mb][nb][e] : acc[mb][nb][e];
                    if (accum) v += cold[mb][nb][e];
                    t.C[row + (size_t)col * t.ldc] = v;
                }
            }
}

// 1 / d to within an ulp in six instructions (the reciprocal refinement div.rn.f64 itself starts with: MUFU.RCP64H seed, two
// Newton steps).  k_panel scales by it instead of dividing: an inlined fp64 division is ~100 SASS instructions with its slow
// path, and the kernel -- straight-line code that each of a few warps executes once -- runs at instruction-fetch speed
// (13 824 -> 12 656 SASS instructions, and warp 0's chain of 32 dependent divisions becomes one of 32 short reciprocals;
// profiles/r02c_factor_levels.txt).  The pivot reciprocals of k_front_fused* use it too.
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = __fma_rn(-d, y, 1.0);
    e = __fma_rn(e, e, e);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-d, y, 1.0);
    return __fma_rn(y, e, y);
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
    __shared__ double rdd[NB];          // 1 / d
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
        // warp 0: lane r holds row r of the lower triangle in registers (statically indexed).  Step c: every lane publishes its
        // entry of column c in shared memory (double-buffered: one __syncwarp per step), reads the pivot and -- as broadcast
        // loads -- the unscaled column, and updates its row: a_r,c2 -= (a_rc / d_c) a_c2,c for c < c2 <= r.  (The first version
        // fetched a_c,c2 from lane c with one 64-bit __shfl_sync per entry: eight SASS instructions each inside this
        // divergent region, 8 600 of the kernel's 12 700 instructions -- and the kernel runs at instruction-fetch speed.)
        const int lane = tid;
        const double tau = pivot_tau(stats);
        __shared__ __align__(16) double colbuf[2][NB];
        double row[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) row[c] = (lane >= c) ? Dg[lane][c] : 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            double* buf = colbuf[c & 1];
            buf[lane] = row[c];
            __syncwarp();
            double dc = buf[c];
            if (c < t.nb && fabs(dc) < tau) {          // static pivot (uniform across the warp and across the CTAs of this block column)
                dc = copysign(tau, dc);
                if (lane == c && t.write_diag) atomicAdd(&stats[4], 1ull);
            }
            const bool act = lane > c;
            const double rdc = fast_rcp(dc);
            const double lr = act ? row[c] * rdc : 0.0;
#pragma unroll
            for (int c2 = c + 1; c2 < NB; ++c2) row[c2] -= lr * buf[c2];
            if (act) Dg[lane][c] = lr;
            if (lane == c) { dd[c] = dc; rdd[c] = rdc; }
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
            if (c < t.nb) L[r + (size_t)(t.k0 + c) * ld] = f[c] * rdd[c];
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
                          const double* __restrict__ B, double sigma, double* __restrict__ L,
                          unsigned long long* __restrict__ stats) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double mx = 0.0;
    for (; s < nnz; s += stride) {
        const int64_t dst = dest[s];
        if (dst >= 0) {
            const double v = B ? A[s] - sigma * B[s] : A[s];
            L[dst] = v;
            mx = fmax(mx, fabs(v));                 // (NaN entries drop out here and surface as bad pivots)
        }
    }
    // max |M_ij|: non-negative doubles order like their bit patterns
    unsigned long long b = (unsigned long long)__double_as_longlong(mx);
#pragma unroll
    for (int o = 16; o; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    if ((threadIdx.x & 31) == 0 && b) atomicMax(&stats[5], b);
}

// ---- accuracy guard: probe right-hand side, residual, correction -------------------------------
__global__ void k_probe_rhs(int64_t N, double* __restrict__ b) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n >= N) return;
    uint64_t z = 0x2545F4914F6CDD1Dull + 0x9E3779B97F4A7C15ull * (uint64_t)(n + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    b[n] = ((double)(z >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
}
// r = b - (ya - sigma * yb)   (ya = A x, yb = B x or nullptr); when nrm != nullptr also max |r|, |x|, |b| -> nrm[0..2]
__global__ void __launch_bounds__(256) k_residual(int64_t N, const double* __restrict__ b, const double* __restrict__ ya,
                                                  const double* __restrict__ yb, double sigma, const double* __restrict__ x,
                                                  double* __restrict__ r, unsigned long long* __restrict__ nrm) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double rv = 0.0, xv = 0.0, bv = 0.0;
    if (n < N) {
        bv = b[n];
        xv = x[n];
        rv = bv - (yb ? ya[n] - sigma * yb[n] : ya[n]);
        r[n] = rv;
    }
    if (nrm) {
        // NaN -> +inf so that a broken solve can never look accurate
        unsigned long long m0 = (unsigned long long)__double_as_longlong(rv == rv ? fabs(rv) : __longlong_as_double(0x7ff0000000000000ll));
        unsigned long long m1 = (unsigned long long)__double_as_longlong(xv == xv ? fabs(xv) : __longlong_as_double(0x7ff0000000000000ll));
        unsigned long long m2 = (unsigned long long)__double_as_longlong(fabs(bv));
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, o));
            m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o));
            m2 = max(m2, __shfl_xor_sync(0xffffffffu, m2, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (m0) atomicMax(&nrm[0], m0);
            if (m1) atomicMax(&nrm[1], m1);
            if (m2) atomicMax(&nrm[2], m2);
        }
    }
}
__global__ void k_add_inplace(int64_t N, double* __restrict__ x, const double* __restrict__ dx) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n < N) x[n] += dx[n];
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
    // the thread's four entries: all loads first, then the four read-modify-writes (distinct targets: rel is injective), each in
    // two passes -- one statement per entry would chain eight dependent global round trips behind stores that might alias
    double* dst[4];
    double v[4], old[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int j = t.tj * 32 + ty + 8 * u;
        dst[u] = nullptr;
        v[u] = 0.0;
        if (j <= i && j < qc) {
            const int rj = rel[j];
            v[u] = Uc[i + (size_t)j * qc];
            dst[u] = rj < pp ? Lp + ri + (size_t)rj * mp : Up + (ri - pp) + (size_t)(rj - pp) * qp;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) old[u] = dst[u] ? *dst[u] : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (dst[u]) *dst[u] = old[u] + v[u];
}

// ---- small fronts: the whole front in one CTA -------------------------------------------------
// Below the top ~10 levels of the tree the fronts are tiny (m <= 256 rows, p <= 64 pivots) and
// there are tens of thousands of them: as tile tasks they cost more in task records, launches
// and tile padding than in arithmetic.  Here one CTA takes one front through every step --
// extend-add of the children, LDL^T of the pivot columns, Schur complement, selective inversion
// -- with the m x p panel resident in shared memory, and writes only what later phases read:
// the update matrix U (parent's extend-add), the inverted panel S and the pivots d (solve sweeps).
//   LDL^T: thread i owns row i of the panel (so the right-looking update needs no atomics and
//   one barrier per pivot: only the pivot column of the p x p block is exchanged, double-buffered).
struct FusedCtx {
    FrontView fv;
    const double* L;     // scattered matrix entries (panel layout), read only
    double* S;
    double* U;
    double* d;
    unsigned long long* stats;
    const int32_t* wsrc0;   // gather maps of the solve sweeps: row of the parent -> position of the
    const int32_t* wsrc1;   // child's boundary entry in W (or -1); reused here for the extend-add
};

__host__ __device__ inline size_t fused_smem_doubles(int m, int p) {
    return 5 * (size_t)((p + 1) & ~1) + 8 + (size_t)(m | 1) * p + (size_t)(m | 1) + 1;
}

template <int PMAX>
__device__ __forceinline__ void front_fused_body(int64_t first, int has_children, const FusedCtx& cx) {
    extern __shared__ __align__(16) double fsm[];
    const int64_t t = first + blockIdx.x;
    const FrontView& fv = cx.fv;
    const int p = fv.p[t], q = fv.q[t], m = p + q;
    const int ld = m | 1, pe = (p + 1) & ~1;
    double* vb = fsm;                       // pivot column exchange, 2 x pe (16-byte aligned halves)
    double* rb = vb + 2 * pe;               // reciprocal pivots, then row staging of the inversion, 2 x pe
    double* dd = rb + 2 * pe;               // pivots
    double* P = dd + pe + 8;                // panel, column-major, ld
    const int tid = threadIdx.x, NT = blockDim.x;
    const double* Lg = cx.L + fv.Loff[t];
    double* Ut = cx.U + fv.Uoff[t];
    // (1)+(2) panel -> shared memory, with the extend-add of the two children folded in as a GATHER:
    // entry (r, c) of this front receives U_child[j(r), j(c)] where j() is the position of the
    // row in the child's boundary list (-1 if absent), read from the sweep gather maps.  Every
    // panel entry and every entry of the update matrix is written exactly once -- no read-modify-
    // write chains through global memory (those serialised the separator fronts).
    int32_t* jm0 = reinterpret_cast<int32_t*>(P + (size_t)ld * p);      // [m] child 0, [m] child 1
    int32_t* jm1 = jm0 + m;
    int qc0 = 0, qc1 = 0;
    const double* U0 = nullptr;
    const double* U1 = nullptr;
    if (has_children) {
        const int64_t c0 = 2 * t, c1 = 2 * t + 1;
        qc0 = fv.q[c0];
        qc1 = fv.q[c1];
        U0 = cx.U + fv.Uoff[c0];
        U1 = cx.U + fv.Uoff[c1];
        if (tid < m) {
            const int64_t w = fv.woff[t] + tid;
            const int32_t s0 = cx.wsrc0[w], s1 = cx.wsrc1[w];
            jm0[tid] = s0 >= 0 ? (int32_t)(s0 - (fv.woff[c0] + fv.p[c0])) : -1;
            jm1[tid] = s1 >= 0 ? (int32_t)(s1 - (fv.woff[c1] + fv.p[c1])) : -1;
        }
        __syncthreads();
    }
    if (tid < m) {
        const double* src = Lg + tid;
        double* dst = P + tid;
        const int jr0 = has_children ? jm0[tid] : -1, jr1 = has_children ? jm1[tid] : -1;
        for (int j = 0; j < p; ++j, src += m, dst += ld) {
            double v = *src;
            if (has_children && j <= tid) {
                const int jc0 = jm0[j], jc1 = jm1[j];
                if (jr0 >= 0 && jc0 >= 0) v += U0[jr0 + (size_t)jc0 * qc0];
                if (jr1 >= 0 && jc1 >= 0) v += U1[jr1 + (size_t)jc1 * qc1];
            }
            *dst = v;
        }
    }
    __syncthreads();
    // (3) LDL^T of the p pivot columns, right-looking, thread i owns row i (no atomics, one barrier per
    // pivot: only the pivot column of the p x p block is exchanged, double-buffered).  The owner of
    // row k+1 has the shortest update of step k, so it also inverts the next pivot while the others
    // finish: the division is off the critical path and rows are scaled by multiplication.
    double* rd = rb;
    // static pivoting: the thread that inverts a pivot also guards it (|pivot| < tau -> +-tau) and records it in
    // dd[]; perturbations are counted per thread and reported once (no atomics inside the elimination loop)
    const double tau = pivot_tau(cx.stats);
    int npert = 0;
    auto guarded = [&](double pv) {
        if (fabs(pv) < tau) {
            pv = copysign(tau, pv);
            ++npert;
        }
        return pv;
    };
    if constexpr (PMAX > 0) {
        // wide pivot blocks (leaf fronts): the row lives in REGISTERS (PMAX columns, statically
        // indexed: the k and j loops are fully unrolled), so one update is one DFMA plus half a
        // broadcast LDS.128; entries above the diagonal or beyond column p are updated with whatever
        // the exchange buffer holds and never read.
        double row[PMAX > 0 ? PMAX : 1];
#pragma unroll
        for (int j = 0; j < PMAX; ++j) row[j] = (tid < m && j < p) ? P[tid + j * ld] : 0.0;
        if (tid == 0 && p > 0) {
            const double pv = guarded(row[0]);
            dd[0] = pv;
            rd[0] = fast_rcp(pv);
        }
#pragma unroll
        for (int k = 0; k < PMAX; ++k) {
            if (k >= p) break;
            double* buf = vb + (k & 1) * pe;
            if (tid >= k && tid < p) buf[tid] = row[k];
            __syncthreads();
            if (tid > k && tid < m) {
                const double li = row[k] * rd[k];
                row[k] = li;
#pragma unroll
                for (int c = 0; c < PMAX / 8; ++c) {
                    if (c * 8 + 7 > k && c * 8 < p) {           // first test static, second uniform
                        const double2* b2 = reinterpret_cast<const double2*>(buf + c * 8);
                        const double2 v0 = b2[0], v1 = b2[1], v2 = b2[2], v3 = b2[3];
                        if (c * 8 + 0 > k) row[c * 8 + 0] -= li * v0.x;
                        if (c * 8 + 1 > k) row[c * 8 + 1] -= li * v0.y;
                        if (c * 8 + 2 > k) row[c * 8 + 2] -= li * v1.x;
                        if (c * 8 + 3 > k) row[c * 8 + 3] -= li * v1.y;
                        if (c * 8 + 4 > k) row[c * 8 + 4] -= li * v2.x;
                        if (c * 8 + 5 > k) row[c * 8 + 5] -= li * v2.y;
                        if (c * 8 + 6 > k) row[c * 8 + 6] -= li * v3.x;
                        if (c * 8 + 7 > k) row[c * 8 + 7] -= li * v3.y;
                    }
                }
                if (k + 1 < PMAX && tid == k + 1 && tid < p) {
                    const double pv = guarded(row[k + 1 < PMAX ? k + 1 : 0]);
                    dd[k + 1] = pv;
                    rd[k + 1] = fast_rcp(pv);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PMAX; ++j)
            if (tid < m && j < p && j <= tid) P[tid + j * ld] = row[j];
    } else {
        // narrow pivot blocks (separator fronts): the panel stays in shared memory -- fewer registers,
        // more resident CTAs, no work on columns that do not exist
        if (tid == 0 && p > 0) {
            const double pv = guarded(P[0]);
            dd[0] = pv;
            rd[0] = fast_rcp(pv);
        }
        for (int k = 0; k < p; ++k) {
            double* buf = vb + (k & 1) * pe;
            if (tid >= k && tid < p) buf[tid] = P[tid + k * ld];
            __syncthreads();
            if (tid > k && tid < m) {
                const double* __restrict__ bj = buf;
                double* __restrict__ pr = P + tid + k * ld;
                const double li = *pr * rd[k];
                *pr = li;
                const int jmax = min(tid, p - 1);
                pr += ld;
                int j = k + 1;
                for (; j + 3 <= jmax; j += 4, pr += 4 * ld) {
                    const double a0 = pr[0], a1 = pr[ld], a2 = pr[2 * ld], a3 = pr[3 * ld];
                    const double b0 = bj[j], b1 = bj[j + 1], b2 = bj[j + 2], b3 = bj[j + 3];
                    pr[0] = a0 - li * b0;
                    pr[ld] = a1 - li * b1;
                    pr[2 * ld] = a2 - li * b2;
                    pr[3 * ld] = a3 - li * b3;
                }
                for (; j <= jmax; ++j, pr += ld) *pr -= li * bj[j];
                if (tid == k + 1 && tid < p) {
                    const double pv = guarded(P[tid + tid * ld]);
                    dd[tid] = pv;
                    rd[tid] = fast_rcp(pv);
                }
            }
        }
    }
    if (npert) atomicAdd(&cx.stats[4], (unsigned long long)npert);
    __syncthreads();
    // pivots + statistics (one set of atomics per warp that holds pivots)
    if (tid < ((p + 31) & ~31)) {
        const bool on = tid < p;
        const double dv = on ? dd[tid] : 1.0;
        const double av = fabs(dv);
        const bool bad = on && (!(av > 0.0) || !isfinite(dv));
        if (on) cx.d[fv.start[t] + tid] = dv;
        const unsigned nneg = __popc(__ballot_sync(0xffffffffu, on && dv < 0.0));
        const unsigned nbad = __popc(__ballot_sync(0xffffffffu, bad));
        unsigned long long lo = (on && !bad) ? (unsigned long long)__double_as_longlong(av) : ~0ull;
        unsigned long long hi = (on && !bad) ? (unsigned long long)__double_as_longlong(av) : 0ull;
        for (int o = 16; o; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((tid & 31) == 0) {
            if (nneg) atomicAdd(&cx.stats[0], (unsigned long long)nneg);
            if (nbad) atomicAdd(&cx.stats[1], (unsigned long long)nbad);
            if (lo != ~0ull) {
                atomicMin(&cx.stats[2], lo);
                atomicMax(&cx.stats[3], hi);
            }
        }
    }
    // (4) Schur complement U -= L21 D L21^T (lower triangle) and (5) X = L11^-1 in place.  They touch
    // disjoint parts of the panel, so with three or more warps the first two invert while the rest
    // update (the inversion has only p-way parallelism).
    const bool split = NT >= 96 && p <= 64;
    auto schur = [&](int t0, int nthr) {
        for (int idx = t0; idx < q * q; idx += nthr) {
            const int a = idx % q, b = idx / q;
            if (a < b) continue;
            const double* la = P + p + a;
            const double* lb = P + p + b;
            // the children's entries are requested before the k loop (read-only in this launch: ld.global.nc, so the loads of
            // one entry do not queue behind the store of the previous one)
            double g = 0.0;
            if (has_children) {
                const int ja0 = jm0[p + a], jb0 = jm0[p + b], ja1 = jm1[p + a], jb1 = jm1[p + b];
                if (ja0 >= 0 && jb0 >= 0) g += __ldg(U0 + ja0 + (size_t)jb0 * qc0);
                if (ja1 >= 0 && jb1 >= 0) g += __ldg(U1 + ja1 + (size_t)jb1 * qc1);
            }
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            int k = 0;
            for (; k + 3 < p; k += 4, la += 4 * ld, lb += 4 * ld) {
                acc0 += la[0] * (dd[k] * lb[0]);
                acc1 += la[ld] * (dd[k + 1] * lb[ld]);
                acc2 += la[2 * ld] * (dd[k + 2] * lb[2 * ld]);
                acc3 += la[3 * ld] * (dd[k + 3] * lb[3 * ld]);
            }
            for (; k < p; ++k, la += ld, lb += ld) acc0 += la[0] * (dd[k] * lb[0]);
            acc0 += acc2;
            acc1 += acc3;
            Ut[a + (size_t)b * q] = g - (acc0 + acc1);
        }
    };
    // row-oriented in-place inversion of the unit lower triangle: row i of X is
    //   X[i][j] = -( L[i][j] + sum_{j<k<i} L[i][k] X[k][j] ),  thread j owns column j (it only ever reads
    //   its own column of X), row i of L is staged so that it can be overwritten.
    auto invert = [&](bool named) {
        for (int i = 1; i < p; ++i) {
            double* row = rb + (i & 1) * pe;
            if (tid < i) row[tid] = P[i + tid * ld];
            if (named) asm volatile("bar.sync 1, 64;" ::: "memory");
            else __syncthreads();
            if (tid < i) {
                const double* __restrict__ xc = P + tid * ld;
                const double* __restrict__ rw = row;
                double s0 = rw[tid], s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int k = tid + 1;
                for (; k + 3 < i; k += 4) {
                    s0 += rw[k] * xc[k];
                    s1 += rw[k + 1] * xc[k + 1];
                    s2 += rw[k + 2] * xc[k + 2];
                    s3 += rw[k + 3] * xc[k + 3];
                }
                for (; k < i; ++k) s0 += rw[k] * xc[k];
                P[i + tid * ld] = -((s0 + s1) + (s2 + s3));
            }
        }
    };
    if (split) {
        if (tid < 64) invert(true);
        else schur(tid - 64, NT - 64);
    } else {
        schur(tid, NT);
        __syncthreads();
        invert(false);
    }
    __syncthreads();
    // (6) inverted panel: S11 = X (unit diagonal), S21 = L21 X
    double* Sg = cx.S + fv.Loff[t];
    if (tid < p) {
        for (int j = 0; j < tid; ++j) Sg[tid + (size_t)j * m] = P[tid + j * ld];
        Sg[tid + (size_t)tid * m] = 1.0;
    }
    for (int idx = tid; idx < q * p; idx += NT) {
        const int a = idx % q, j = idx / q;
        const double* la = P + p + a + j * ld;
        const double* xj = P + j * ld;              // X[k][j], k > j; X[j][j] = 1
        double acc0 = la[0], acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        la += ld;
        int k = j + 1;
        for (; k + 3 < p; k += 4, la += 4 * ld) {
            acc0 += la[0] * xj[k];
            acc1 += la[ld] * xj[k + 1];
            acc2 += la[2 * ld] * xj[k + 2];
            acc3 += la[3 * ld] * xj[k + 3];
        }
        for (; k < p; ++k, la += ld) acc0 += la[0] * xj[k];
        Sg[p + a + (size_t)j * m] = (acc0 + acc1) + (acc2 + acc3);
    }
}

// two entry points over one body: the register-resident variant is capped at 160 registers (ptxas otherwise
// takes 230+ and loses a resident CTA per SM on the leaf level); the shared-memory variant needs ~40
template <int PMAX>
__global__ void __maxnreg__(160) k_front_fused_regs(int64_t first, int has_children, FusedCtx cx) {
    front_fused_body<PMAX>(first, has_children, cx);
}
__global__ void k_front_fused(int64_t first, int has_children, FusedCtx cx) { front_fused_body<0>(first, has_children, cx); }

// ---- solve sweeps -----------------------------------------------------------------------------
// x = M^-1 b with the selectively inverted panels S:  forward  y_own = S11 w_own, w_bnd -= S21 w_own
// (w = b restricted to the front + the children's boundary remainders), backward
// x_own = S11^T (D^-1 y_own) - S21^T x_bnd.  One launch per tree level and direction; nothing else:
//   * every front has ONE 48-byte record (no chain of per-field lookups);
//   * the children's contributions are fetched through gather maps built once per mesh
//     (wsrc0/wsrc1[row of the parent] = position of the child's remainder in W, or -1), so a
//     front needs no child metadata and no separate gather kernel, and the sum order is fixed;
//   * the matrix loads do not depend on the vectors, so every thread first puts a batch of S
//     loads in flight and only then resolves the vector dependency (the kernels are bound by
//     the latency of that chain, not by arithmetic);
//   * the row / column permutation and the D^-1 scaling are folded into the first and last touch.
// Programmatic dependent launch: consecutive sweep kernels form a chain (level l needs level l+1), but
// the matrix loads of a level do not depend on the previous one.  A kernel lets its successor start
// as soon as all of its own CTAs are resident (launch_dependents), the successor puts its panel loads
// in flight and only then waits for the predecessor's vector data (wait): ramp-up and tail of these
// 20-40 us launches overlap.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Per-front completion flags replace the level-wide wait between consecutive sweep kernels: a front
// starts as soon as ITS OWN inputs are complete (forward: its two children; backward: its parent and
// its own forward pass), so the ramp and tail of a level overlap with its neighbours instead of
// adding up.  Safe against deadlock because a programmatically launched grid only becomes resident
// once every CTA of its predecessor has started (each kernel triggers at its first instruction): a
// spinning CTA always waits for CTAs that are already running.  The spin is bounded so that a logic
// error can never hang the GPU (the result is then wrong and the parity tests say so).
__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void flag_wait(const int32_t* f, int need, int32_t* abort_flag) {
    for (int it = 0; it < (1 << 21); ++it) {
        if (ld_acquire_gpu(f) >= need) return;
        if ((it & 1023) == 1023 && ld_acquire_gpu(abort_flag)) return;
    }
    atomicExch(abort_flag, 1);      // every later wait returns at once; ws_solver_solve's caller sees wrong numbers
}
// every thread of the CTA has issued its global writes; publish them and count this CTA in
__device__ __forceinline__ void flag_signal(int32_t* f) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(f, 1);
    }
}

struct FrontRec {
    int64_t Loff, start, woff, bndoff;
    int32_t p, q, pad0, pad1;       // pad0 / pad1: CTAs the forward / backward sweep spends on this front
    int32_t cf0, cf1, pb, pad2;     // the same counts of the two children (forward) and of the parent (backward): a waiter
                                    // finds what it has to wait for in its own record, not behind another dependent load
};

constexpr int FRONT_MAX_M = 1024;   // fronts up to this size are handled by one CTA

struct SolveCtx {
    const FrontRec* recs;
    const double* S;
    const int32_t* perm;
    const int32_t* wsrc0;
    const int32_t* wsrc1;
    const int32_t* bnd;
    const double* dinv;
    double* W;      // boundary remainders of every front (forward)
    double* yd;     // D^-1 y, permuted order
    double* x;      // solution, permuted order
    int64_t sW;     // stride between right-hand sides in W
    int64_t sN;     // stride between right-hand sides in yd / x / b / xout (= N)
    int32_t* ffl;   // per front: CTAs of the forward sweep that have finished (nullptr: level-wide waits)
    int32_t* bfl;   // the same for the backward sweep
};

// forward: wait for the two children of front t (their CTA counts are in their records)
__device__ __forceinline__ void fwd_wait_children(const SolveCtx& cx, int64_t t, const FrontRec& R) {
    if (threadIdx.x == 0) {
        flag_wait(cx.ffl + 2 * t, R.cf0, cx.ffl);          // slot 0 (front ids start at 1) = abort flag
        flag_wait(cx.ffl + 2 * t + 1, R.cf1, cx.ffl);
    }
    __syncthreads();
}
// backward: wait for the front's own forward pass and for its parent's backward pass
__device__ __forceinline__ void bwd_wait_parent(const SolveCtx& cx, int64_t t, const FrontRec& R) {
    if (threadIdx.x == 0) {
        flag_wait(cx.ffl + t, R.pad0, cx.ffl);
        if (t > 1) flag_wait(cx.bfl + (t >> 1), R.pb, cx.ffl);
    }
    __syncthreads();
}

// NR right-hand sides share every load of the inverted panels: a sweep is bound by the panel bytes,
// so a second vector costs its own (small) vector traffic and arithmetic only.
// the static part of the gather (permutation / child positions): can be loaded before the dependency is resolved
struct GatherIdx {
    int32_t pi, s0, s1;
};
__device__ __forceinline__ GatherIdx fwd_gather_idx(const SolveCtx& cx, const FrontRec& R, int i, int has_children) {
    GatherIdx g;
    g.pi = (i < R.p) ? cx.perm[R.start + i] : -1;
    g.s0 = g.s1 = -1;
    if (has_children) {
        g.s0 = cx.wsrc0[R.woff + i];
        g.s1 = cx.wsrc1[R.woff + i];
    }
    return g;
}
template <int NR>
__device__ __forceinline__ void fwd_gather_val(const SolveCtx& cx, const GatherIdx& g, const double* __restrict__ b, double (&v)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double t = g.pi >= 0 ? b[g.pi + r * cx.sN] : 0.0;
        if (g.s0 >= 0) t += cx.W[g.s0 + r * cx.sW];
        if (g.s1 >= 0) t += cx.W[g.s1 + r * cx.sW];
        v[r] = t;
    }
}
template <int NR>
__device__ __forceinline__ void fwd_gather(const SolveCtx& cx, const FrontRec& R, int i, int has_children,
                                           const double* __restrict__ b, double (&v)[NR]) {
    const int32_t pi = (i < R.p) ? cx.perm[R.start + i] : -1;
    int s0 = -1, s1 = -1;
    if (has_children) {
        s0 = cx.wsrc0[R.woff + i];
        s1 = cx.wsrc1[R.woff + i];
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double t = pi >= 0 ? b[pi + r * cx.sN] : 0.0;
        if (s0 >= 0) t += cx.W[s0 + r * cx.sW];
        if (s1 >= 0) t += cx.W[s1 + r * cx.sW];
        v[r] = t;
    }
}

// one CTA per front, one thread per front row (blockDim.x >= m, multiple of 32, <= MAXT).
// The k loop is double-buffered: KB loads of the next batch are in flight while the current
// batch is multiplied.  (MAXT only sets the register budget: with a 1024-thread bound ptxas has
// 64 registers and serialises the batch, so small fronts get their own instantiation.)
template <int MAXT, int KB, int NR>
__global__ void __launch_bounds__(MAXT) k_fwd_front(int64_t first, int has_children, SolveCtx cx,
                                                    const double* __restrict__ b, int pstride) {
    extern __shared__ double sw[];          // [NR][pstride]
    pdl_launch_dependents();
    const FrontRec R = cx.recs[first + blockIdx.x];
    const int p = R.p, m = p + R.q, r = threadIdx.x;
    const bool act = r < m;
    const int kend = act ? (r < p ? r + 1 : p) : 0;
    const double* Sr = cx.S + R.Loff + r;
    const size_t ld = (size_t)m;
    double a[KB], c[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) a[j] = (j < kend) ? __ldg(Sr + (size_t)j * ld) : 0.0;
#pragma unroll
    for (int j = 0; j < KB; ++j) c[j] = (KB + j < kend) ? __ldg(Sr + (size_t)(KB + j) * ld) : 0.0;
    double wv[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) wv[q] = 0.0;
    GatherIdx gi{-1, -1, -1};
    double di = 0.0;
    if (act) {
        gi = fwd_gather_idx(cx, R, r, has_children);
        if (r < p) di = cx.dinv[R.start + r];
    }
    if (cx.ffl && has_children) fwd_wait_children(cx, first + blockIdx.x, R);
    else pdl_wait();
    if (act) {
        fwd_gather_val<NR>(cx, gi, b, wv);
        if (r < p) {
#pragma unroll
            for (int q = 0; q < NR; ++q) sw[q * pstride + r] = wv[q];
        }
    }
    __syncthreads();
    double acc0[NR], acc1[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) acc0[q] = acc1[q] = 0.0;
    for (int k0 = 0; k0 < kend; k0 += 2 * KB) {
#pragma unroll
        for (int j = 0; j < KB; j += 2) {
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                if (k0 + j < kend) acc0[q] += a[j] * sw[q * pstride + k0 + j];
                if (k0 + j + 1 < kend) acc1[q] += a[j + 1] * sw[q * pstride + k0 + j + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) a[j] = (k0 + 2 * KB + j < kend) ? __ldg(Sr + (size_t)(k0 + 2 * KB + j) * ld) : 0.0;
#pragma unroll
        for (int j = 0; j < KB; j += 2) {
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                if (k0 + KB + j < kend) acc0[q] += c[j] * sw[q * pstride + k0 + KB + j];
                if (k0 + KB + j + 1 < kend) acc1[q] += c[j + 1] * sw[q * pstride + k0 + KB + j + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) c[j] = (k0 + 3 * KB + j < kend) ? __ldg(Sr + (size_t)(k0 + 3 * KB + j) * ld) : 0.0;
    }
    if (act) {
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const double acc = acc0[q] + acc1[q];
            if (r < p) cx.yd[R.start + r + q * cx.sN] = acc * di;
            else cx.W[R.woff + r + q * cx.sW] = wv[q] - acc;
        }
    }
    if (cx.ffl) flag_signal(cx.ffl + first + blockIdx.x);
}

// large fronts: 32 rows x KL k-lanes per CTA (task = row tile of one front)
struct FwdTile {
    int32_t front, r0, nr, pad;
};
template <int KL, int NR>
__global__ void __launch_bounds__(32 * KL) k_fwd_tile(const FwdTile* __restrict__ tasks, int has_children, SolveCtx cx,
                                                      const double* __restrict__ b, int pstride) {
    extern __shared__ double sm[];
    pdl_launch_dependents();
    const FwdTile t = tasks[blockIdx.x];
    const FrontRec R = cx.recs[t.front];
    const int p = R.p, m = p + R.q;
    double* sw = sm;                       // [NR][pstride]
    const int lr = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int r = t.r0 + lr;
    const bool valid = lr < t.nr;
    const int kend_tile = min(p, (t.r0 < p) ? t.r0 + t.nr : p);
    double* red = sm + (size_t)NR * pstride;        // [NR][KL][33]
    const int kend = valid ? ((r < p) ? min(kend_tile, r + 1) : p) : 0;
    const double* Sr = cx.S + R.Loff + r;
    const size_t ld = (size_t)m;
    constexpr int NL = 8;                  // loads in flight per thread
    double a[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) a[j] = (kl + j * KL < kend) ? __ldg(Sr + (size_t)(kl + j * KL) * ld) : 0.0;
    if (cx.ffl && has_children) fwd_wait_children(cx, t.front, R);
    else pdl_wait();
    for (int k = threadIdx.x; k < kend_tile; k += 32 * KL) {
        double g[NR];
        fwd_gather<NR>(cx, R, k, has_children, b, g);
#pragma unroll
        for (int q = 0; q < NR; ++q) sw[q * pstride + k] = g[q];
    }
    double wv[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) wv[q] = 0.0;
    if (kl == 0 && valid && r >= p) fwd_gather<NR>(cx, R, r, has_children, b, wv);
    __syncthreads();
    double acc[NR], acc1[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) acc[q] = acc1[q] = 0.0;
    for (int k = kl; k < kend; k += NL * KL) {
        double c[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) c[j] = a[j];
#pragma unroll
        for (int j = 0; j < NL; ++j)
            a[j] = (k + (NL + j) * KL < kend) ? __ldg(Sr + (size_t)(k + (NL + j) * KL) * ld) : 0.0;
#pragma unroll
        for (int j = 0; j < NL; j += 2) {
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                if (k + j * KL < kend) acc[q] += c[j] * sw[q * pstride + k + j * KL];
                if (k + (j + 1) * KL < kend) acc1[q] += c[j + 1] * sw[q * pstride + k + (j + 1) * KL];
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NR; ++q) red[(q * KL + kl) * 33 + lr] = acc[q] + acc1[q];
    __syncthreads();
    if (kl == 0 && valid) {
        const double di = r < p ? cx.dinv[R.start + r] : 0.0;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < KL; ++j) sacc += red[(q * KL + j) * 33 + lr];
            if (r < p) cx.yd[R.start + r + q * cx.sN] = sacc * di;
            else cx.W[R.woff + r + q * cx.sW] = wv[q] - sacc;
        }
    }
    if (cx.ffl) flag_signal(cx.ffl + t.front);
}

// backward, one CTA per front: warp w owns columns w, w+NW, ...; BWD_CB columns x BWD_RJ row groups
// of loads in flight per lane; g = [D^-1 y_own ; -x_bnd] is staged in shared memory.
template <int MAXT, int BWD_CB, int BWD_RJ, int NR>
__global__ void __launch_bounds__(MAXT) k_bwd_front(int64_t first, SolveCtx cx, double* __restrict__ xout, int mstride) {
    extern __shared__ double sg[];          // [NR][mstride]
    pdl_launch_dependents();
    const FrontRec R = cx.recs[first + blockIdx.x];
    const int p = R.p, m = p + R.q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
    const double* Sf = cx.S + R.Loff;
    double a[BWD_CB][BWD_RJ];
#pragma unroll
    for (int i = 0; i < BWD_CB; ++i) {
        const int c = warp + NW * i;
#pragma unroll
        for (int j = 0; j < BWD_RJ; ++j) {
            const int r = c + lane + 32 * j;
            a[i][j] = (c < p && r < m) ? __ldg(Sf + r + (size_t)c * m) : 0.0;
        }
    }
    int64_t src0 = 0;
    if ((int)threadIdx.x < m) src0 = ((int)threadIdx.x < p) ? R.start + threadIdx.x : (int64_t)cx.bnd[R.bndoff + threadIdx.x - p];
    if (cx.bfl) bwd_wait_parent(cx, first + blockIdx.x, R);
    else pdl_wait();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int64_t src = i == (int)threadIdx.x ? src0 : ((i < p) ? R.start + i : (int64_t)cx.bnd[R.bndoff + i - p]);
#pragma unroll
        for (int q = 0; q < NR; ++q) sg[q * mstride + i] = (i < p) ? cx.yd[src + q * cx.sN] : -cx.x[src + q * cx.sN];
    }
    __syncthreads();
    for (int cb = warp; cb < p; cb += NW * BWD_CB) {
        double acc[BWD_CB][NR];
#pragma unroll
        for (int i = 0; i < BWD_CB; ++i) {
            const int c = cb + NW * i;
#pragma unroll
            for (int q = 0; q < NR; ++q) acc[i][q] = 0.0;
#pragma unroll
            for (int j = 0; j < BWD_RJ; ++j) {
                const int r = c + lane + 32 * j;
                if (c < p && r < m) {
#pragma unroll
                    for (int q = 0; q < NR; ++q) acc[i][q] += a[i][j] * sg[q * mstride + r];
                }
            }
        }
        // remaining row groups of this batch
        const int rmax = m - cb;                    // rows below the first column of the batch
        for (int j0 = BWD_RJ; j0 * 32 < rmax; j0 += BWD_RJ) {
#pragma unroll
            for (int i = 0; i < BWD_CB; ++i) {
                const int c = cb + NW * i;
#pragma unroll
                for (int j = 0; j < BWD_RJ; ++j) {
                    const int r = c + lane + 32 * (j0 + j);
                    a[i][j] = (c < p && r < m) ? __ldg(Sf + r + (size_t)c * m) : 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < BWD_CB; ++i) {
                const int c = cb + NW * i;
#pragma unroll
                for (int j = 0; j < BWD_RJ; ++j) {
                    const int r = c + lane + 32 * (j0 + j);
                    if (c < p && r < m) {
#pragma unroll
                        for (int q = 0; q < NR; ++q) acc[i][q] += a[i][j] * sg[q * mstride + r];
                    }
                }
            }
        }
        // prefetch the next batch while this one is reduced
        const int nb = cb + NW * BWD_CB;
#pragma unroll
        for (int i = 0; i < BWD_CB; ++i) {
            const int c = nb + NW * i;
#pragma unroll
            for (int j = 0; j < BWD_RJ; ++j) {
                const int r = c + lane + 32 * j;
                a[i][j] = (c < p && r < m) ? __ldg(Sf + r + (size_t)c * m) : 0.0;
            }
        }
#pragma unroll
        for (int i = 0; i < BWD_CB; ++i) {
            const int c = cb + NW * i;
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                double s = acc[i][q];
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && c < p) {
                    cx.x[R.start + c + q * cx.sN] = s;
                    xout[cx.perm[R.start + c] + q * cx.sN] = s;
                }
            }
        }
    }
    if (cx.bfl) flag_signal(cx.bfl + first + blockIdx.x);
}

// large fronts: GEMVT_COLS columns per CTA, WPC warps share a column (task = column tile)
struct BwdTile {
    int32_t front, c0, nc, pad;
};
template <int WPC, int NR>
__global__ void __launch_bounds__(32 * GEMVT_COLS * WPC) k_bwd_tile(const BwdTile* __restrict__ tasks, SolveCtx cx,
                                                                    double* __restrict__ xout, int mstride) {
    extern __shared__ double sm[];
    pdl_launch_dependents();
    const BwdTile t = tasks[blockIdx.x];
    const FrontRec R = cx.recs[t.front];
    const int p = R.p, m = p + R.q;
    double* sg = sm;                                   // [NR][mstride]: g[c0 .. m)
    double* red = sm + (size_t)NR * mstride;           // [NR][GEMVT_COLS][WPC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wc = warp / WPC, part = warp % WPC;
    const int c = t.c0 + wc;
    const bool on = wc < t.nc;
    const double* Sc = cx.S + R.Loff + (size_t)c * m;
    constexpr int step = 32 * WPC;
    constexpr int NL = 8;                              // loads in flight per lane
    const int rb = c + part * 32 + lane;               // first row of this lane
    double a[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) a[j] = (on && rb + j * step < m) ? __ldg(Sc + rb + j * step) : 0.0;
    if (cx.bfl) bwd_wait_parent(cx, t.front, R);
    else pdl_wait();
    for (int i = t.c0 + threadIdx.x; i < m; i += blockDim.x) {
        const int64_t src = (i < p) ? R.start + i : (int64_t)cx.bnd[R.bndoff + i - p];
#pragma unroll
        for (int q = 0; q < NR; ++q) sg[q * mstride + i - t.c0] = (i < p) ? cx.yd[src + q * cx.sN] : -cx.x[src + q * cx.sN];
    }
    __syncthreads();
    double acc[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) acc[q] = 0.0;
    if (on) {
        double a0[NR], a1[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) a0[q] = a1[q] = 0.0;
        for (int r = rb; r < m; r += NL * step) {
            double cbuf[NL];
#pragma unroll
            for (int j = 0; j < NL; ++j) cbuf[j] = a[j];
#pragma unroll
            for (int j = 0; j < NL; ++j) a[j] = (r + (NL + j) * step < m) ? __ldg(Sc + r + (NL + j) * step) : 0.0;
#pragma unroll
            for (int j = 0; j < NL; j += 2) {
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    if (r + j * step < m) a0[q] += cbuf[j] * sg[q * mstride + r + j * step - t.c0];
                    if (r + (j + 1) * step < m) a1[q] += cbuf[j + 1] * sg[q * mstride + r + (j + 1) * step - t.c0];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            double s = a0[q] + a1[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            acc[q] = s;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NR; ++q) red[(q * GEMVT_COLS + wc) * WPC + part] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < t.nc) {
        const int cc = t.c0 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            double sacc = 0.0;
#pragma unroll
            for (int j = 0; j < WPC; ++j) sacc += red[(q * GEMVT_COLS + threadIdx.x) * WPC + j];
            cx.x[R.start + cc + q * cx.sN] = sacc;
            xout[cx.perm[R.start + cc] + q * cx.sN] = sacc;
        }
    }
    if (cx.bfl) flag_signal(cx.bfl + t.front);
}

// gather maps: for child c of front t = c/2 and boundary entry j of c:
//   wsrc{c&1}[woff[t] + rel_c[j]] = woff[c] + p_c + j
__global__ void k_build_wsrc(int64_t nfronts, const FrontRec* __restrict__ recs, const int32_t* __restrict__ rel,
                             int32_t* __restrict__ wsrc0, int32_t* __restrict__ wsrc1) {
    const int64_t c = 2 + blockIdx.x;
    if (c >= nfronts) return;
    const FrontRec Rc = recs[c], Rp = recs[c >> 1];
    int32_t* dst = (c & 1) ? wsrc1 : wsrc0;
    for (int j = threadIdx.x; j < Rc.q; j += blockDim.x)
        dst[Rp.woff + rel[Rc.bndoff + j]] = (int32_t)(Rc.woff + Rc.p + j);
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

template <class T>
static void clone(DevBuf<T>& dst, const DevBuf<T>& src, size_t n, cudaStream_t st) {
    dst.alloc(n, st);
    if (n) WS_CUDA(cudaMemcpyAsync(dst.get(), src.get(), n * sizeof(T), cudaMemcpyDeviceToDevice, st));
}

static FrontView view(const ws_solver* s) {
    return FrontView{s->fr_p.get(), s->fr_q.get(), s->fr_start.get(), s->fr_Loff.get(), s->fr_Uoff.get(),
                     s->fr_woff.get(), s->fr_bndoff.get(), s->bnd.get(), s->rel.get()};
}

constexpr int OUTER_NB = 4 * NB;                  // outer block of the right-looking top levels
constexpr int FUSED_MAX_M = 256;                  // one thread per front row
constexpr size_t FUSED_MAX_SMEM = 72 * 1024;     // >= 3 resident CTAs per SM
static bool fused_enabled() {
    static const bool on = !(getenv("WS_NO_FUSED") && atoi(getenv("WS_NO_FUSED")));
    return on;
}

// per-level shape and kernel choice (cheap: one pass over the per-front sizes); done at creation
static void classify_levels(ws_solver* s) {
    const Symbolic& Y = s->sym;
    const int depth = Y.depth;
    s->levels.assign(depth + 1, LevelPlan());
    for (int l = depth; l >= 0; --l) {
        LevelPlan& LP = s->levels[l];
        const int64_t f0 = (int64_t)1 << l, f1 = (int64_t)1 << (l + 1);
        LP.first_front = f0;
        LP.nfronts = f1 - f0;
        int maxp = 0, maxm = 0;
        for (int64_t t = f0; t < f1; ++t) {
            maxp = std::max(maxp, (int)Y.p[t]);
            maxm = std::max(maxm, (int)(Y.p[t] + Y.q[t]));
        }
        const int nkb = (maxp + NB - 1) / NB;
        LP.upd.assign(nkb, Range());
        LP.pan.assign(nkb, Range());
        LP.trail.assign(nkb, Range());
        // One CTA per front means one THREAD per front row walking all p columns of its row: p / 8 dependent round trips.
        // Near the root of a small problem (a ~50k-triangle mesh of a sweep: root front 600 x 300) that chain, not bytes,
        // is the level's cost, and there are too few fronts to hide it; such levels (at most WS_TILE_TOP = 32 fronts, at least
        // 64 pivots) take the tile kernels, which split a row's k range over the lanes of a warp / 8 warps: one solve of
        // N = 100k goes from 0.297 to 0.138 ms (profiles/r02c_small_problem_top_levels.txt).  Levels whose fronts exceed one
        // CTA are tiled as before.
        static const int tile_top = getenv("WS_TILE_TOP") ? atoi(getenv("WS_TILE_TOP")) : 32;
        static const int tile_minp = getenv("WS_TILE_MINP") ? atoi(getenv("WS_TILE_MINP")) : 64;
        const bool retile = maxm <= FRONT_MAX_M && (f1 - f0) <= tile_top && maxp >= tile_minp;
        LP.small = maxm <= FRONT_MAX_M && !retile;
        LP.narrow_tiles = retile;
        LP.nt = std::max(32, (maxm + 31) & ~31);
        LP.maxp = maxp;
        LP.maxm = maxm;
        // few large fronts: a left-looking block-column update has too few tiles to fill the GPU
        // (and a long serial k loop); update the whole trailing panel after every block column.
        LP.right_looking = (f1 - f0) <= 64 && maxp > 4 * NB;
        LP.fused = fused_enabled() && maxm <= FUSED_MAX_M && maxp <= 64 &&
                   fused_smem_doubles(maxm, maxp) * sizeof(double) <= FUSED_MAX_SMEM;
    }
}

// Tile-task records of the fronts that are not handled by k_front_fused.  Emitted by the first
// factorisation AFTER it has launched the fused levels, so that the host loop (8 ms at 1M triangles)
// runs while the GPU works on the small fronts; uploads are asynchronous (host copies live in *s).
static void build_plan(ws_solver* s, cudaStream_t st) {
    const Symbolic& Y = s->sym;
    const int depth = Y.depth;
    std::vector<GemmTask>& G = s->h_gemm;
    std::vector<PanelTask>& P = s->h_panel;
    std::vector<ExtTask>& E = s->h_ext;
    std::vector<InvTask>& I = s->h_inv;
    std::vector<FwdTile>& F = s->h_fwd;
    std::vector<BwdTile>& Bk = s->h_bwd;
    G.clear(); P.clear(); E.clear(); I.clear(); F.clear(); Bk.clear();
    double* Lb = s->L.get();
    double* Sb = s->S.get();
    double* Tb = s->T.get();
    double* Ub = s->U.get();
    double* db = s->d.get();
    for (int l = depth; l >= 0; --l) {
        LevelPlan& LP = s->levels[l];
        const int64_t f0 = (int64_t)1 << l, f1 = (int64_t)1 << (l + 1);
        const int maxp = LP.maxp;
        const int nkb = (maxp + NB - 1) / NB;
        // extend-add of the children of this level's fronts (left children first, then right)
        if (l < depth && !LP.fused) {
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
        for (int kb = 0; kb < (LP.fused ? 0 : nkb); ++kb) {
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
                // two-level blocking: after block column kb only the rest of its OUTER block (OB columns)
                // is updated (rank NB, cheap: m x <= OB-NB entries); the trailing matrix beyond the outer
                // block gets one rank-OB update when the outer block is complete -- a quarter of the
                // passes over the trailing matrix and four times the arithmetic intensity per pass.
                for (int64_t t = f0; t < f1; ++t) {
                    const int p = Y.p[t], m = p + Y.q[t];
                    if (p <= k0) continue;
                    const int nbk = std::min(NB, p - k0);
                    const int c0 = k0 + nbk;           // first trailing row / column
                    const int J0 = (k0 / OUTER_NB) * OUTER_NB, Jend = std::min(J0 + OUTER_NB, p);
                    double* Lt = Lb + Y.Loff[t];
                    for (int i0 = c0; i0 < m; i0 += GT)
                        for (int j0 = c0; j0 <= i0 && j0 < Jend; j0 += GT) {
                            GemmTask g{};
                            g.A = Lt + i0 + (size_t)k0 * m;
                            g.B = Lt + j0 + (size_t)k0 * m;
                            g.C = Lt + i0 + (size_t)j0 * m;
                            g.dscale = db + Y.start[t] + k0;
                            g.lda = g.ldb = g.ldc = m;
                            g.m = std::min(GT, m - i0);
                            g.n = std::min(GT, Jend - j0);     // never beyond the outer block
                            g.k = nbk;
                            g.flags = GF_ACCUM | GF_NEG;
                            G.push_back(g);
                        }
                    if (c0 >= Jend && Jend < p)
                        for (int i0 = Jend; i0 < m; i0 += GT)
                            for (int j0 = Jend; j0 <= i0 && j0 < p; j0 += GT) {
                                GemmTask g{};
                                g.A = Lt + i0 + (size_t)J0 * m;
                                g.B = Lt + j0 + (size_t)J0 * m;
                                g.C = Lt + i0 + (size_t)j0 * m;
                                g.dscale = db + Y.start[t] + J0;
                                g.lda = g.ldb = g.ldc = m;
                                g.m = std::min(GT, m - i0);
                                g.n = std::min(GT, p - j0);
                                g.k = Jend - J0;
                                g.flags = GF_ACCUM | GF_NEG;
                                G.push_back(g);
                            }
                }
            LP.trail[kb].cnt = (int64_t)G.size() - LP.trail[kb].off;
        }
        // Schur complement U -= L21 D L21^T (lower tiles)
        LP.syrk.off = (int64_t)G.size();
        for (int64_t t = f0; t < f1 && !LP.fused; ++t) {
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
            // (a front without own unknowns still forwards its children's remainders)
            for (int r0 = 0; r0 < m; r0 += GEMV_ROWS) {
                // a tile must not straddle the own / boundary split (different k extents)
                int nr = std::min(GEMV_ROWS, m - r0);
                if (r0 < p && r0 + nr > p) nr = p - r0;
                F.push_back(FwdTile{(int32_t)t, r0, nr, 0});
                if (nr < GEMV_ROWS && r0 + nr < m)       // remainder of a straddling tile
                    F.push_back(FwdTile{(int32_t)t, r0 + nr, std::min(GEMV_ROWS - nr, m - (r0 + nr)), 0});
            }
            for (int c0 = 0; c0 < p; c0 += GEMVT_COLS)
                Bk.push_back(BwdTile{(int32_t)t, c0, std::min(GEMVT_COLS, p - c0), 0});
        }
        LP.fwd.cnt = (int64_t)F.size() - LP.fwd.off;
        LP.bwd.cnt = (int64_t)Bk.size() - LP.bwd.off;
        {
            const int64_t mid = f0 + (f1 - f0) / 2;
            LP.fwd_mid = 0;
            LP.bwd_mid = 0;
            for (int64_t i = LP.fwd.off; i < (int64_t)F.size() && F[i].front < mid; ++i) ++LP.fwd_mid;
            for (int64_t i = LP.bwd.off; i < (int64_t)Bk.size() && Bk[i].front < mid; ++i) ++LP.bwd_mid;
        }
    }
    // ---- selective inversion (all fronts that are not handled by k_front_fused, independent of the tree)
    auto tiled = [&](int64_t t) { return !s->levels[63 - __builtin_clzll((unsigned long long)t)].fused; };
    s->inv0.off = 0;
    for (int64_t t = 1; t < Y.nfronts; ++t) {
        if (!tiled(t)) continue;
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
            if (!tiled(t)) continue;
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
            if (!tiled(t)) continue;
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
        if (!p || !q || !tiled(t)) continue;
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
    const bool trace = getenv("WS_TRACE") != nullptr;
    const double tg = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    upload(s->gemm_tasks, G, st);
    upload(s->panel_tasks, P, st);
    upload(s->ext_tasks, E, st);
    upload(s->inv_tasks, I, st);
    upload(s->fwd_tiles, F, st);
    upload(s->bwd_tiles, Bk, st);
    s->plan_built = true;
    if (trace)
        for (int l = 0; l <= depth; ++l)
            fprintf(stderr, "[build_plan] level %2d fronts %6lld maxp %4d maxm %4d %s nt %d\n", l, (long long)s->levels[l].nfronts,
                    s->levels[l].maxp, s->levels[l].maxm, s->levels[l].fused ? "fused" : (s->levels[l].small ? "front" : "tile"), s->levels[l].nt);
    if (trace)
        fprintf(stderr, "[build_plan] tasks: gemm %zu panel %zu ext %zu inv %zu fwd %zu bwd %zu; upload %.1f ms\n", G.size(),
                P.size(), E.size(), I.size(), F.size(), Bk.size(),
                (std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - tg) * 1e3);
    s->bytes += (int64_t)(G.size() * sizeof(GemmTask) + P.size() * sizeof(PanelTask) + E.size() * sizeof(ExtTask) +
                          I.size() * sizeof(InvTask) + F.size() * sizeof(FwdTile) + Bk.size() * sizeof(BwdTile));
}

static void launch_gemm(const ws_solver* s, const Range& r, bool narrow, cudaStream_t st) {
    if (!r.cnt) return;
    static PerDeviceOnce configure;
    configure([] {
        WS_CUDA(cudaFuncSetAttribute(k_gemm_tiles<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<32>()));
        WS_CUDA(cudaFuncSetAttribute(k_gemm_tiles<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()));
    });
    if (narrow) k_gemm_tiles<32><<<(unsigned)r.cnt, 128, gemm_smem_bytes<32>(), st>>>(s->gemm_tasks.get() + r.off);
    else k_gemm_tiles<64><<<(unsigned)r.cnt, 128, gemm_smem_bytes<64>(), st>>>(s->gemm_tasks.get() + r.off);
    count_launch(1);
}

// ---- one-CTA-per-front sweep kernels: variant selection ----------------------------------------
// These CTAs are small (one thread per front row) and short-lived, so the number of resident
// CTAs -- bounded by registers -- matters more than the depth of the per-thread load pipeline:
// KB (forward: loads in flight per thread = 2*KB) and CB x RJ (backward: columns x row groups in
// flight per lane) default to small values.  WS_FWD_KB[_LEAF], WS_BWD_CB[_LEAF], WS_BWD_RJ[_LEAF]
// override them for experiments (tools/sweep_variants.py).
// launch with the programmatic-stream-serialisation attribute (the kernel may start before its
// predecessor in the stream has finished; it synchronises itself with griddepcontrol.wait)
static thread_local bool g_pdl_on = true;        // set by sweep_launches from the solver's overlap mode
template <class... KArgs, class... Args>
static void launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
    static const bool pdl_env = !(getenv("WS_NO_PDL") && atoi(getenv("WS_NO_PDL")));
    const bool pdl = pdl_env && g_pdl_on;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    WS_CUDA(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

struct SweepTuning {
    int kb, kb_leaf, cb, cb_leaf, rj, rj_leaf;
    int kb_mid, cb_mid, rj_mid, mid_fronts;   // levels with at most mid_fronts fronts (about one wave of CTAs)
};
static SweepTuning sweep_tuning_read();
static const SweepTuning& sweep_tuning() {
    static SweepTuning t = sweep_tuning_read();
    static const bool reload = getenv("WS_TUNE_RELOAD") && atoi(getenv("WS_TUNE_RELOAD"));   // experiments: re-read per solve
    if (reload) t = sweep_tuning_read();
    return t;
}
static SweepTuning sweep_tuning_read() {
    {
        auto env = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
        SweepTuning v;
        v.kb = env("WS_FWD_KB", 4);
        v.kb_leaf = env("WS_FWD_KB_LEAF", 8);
        v.cb = env("WS_BWD_CB", 1);
        v.cb_leaf = env("WS_BWD_CB_LEAF", 2);
        v.rj = env("WS_BWD_RJ", 4);
        v.rj_leaf = env("WS_BWD_RJ_LEAF", 2);
        v.kb_mid = env("WS_FWD_KB_MID", v.kb);
        v.cb_mid = env("WS_BWD_CB_MID", v.cb);
        v.rj_mid = env("WS_BWD_RJ_MID", v.rj);
        v.mid_fronts = env("WS_MID_FRONTS", 0);
        return v;
    }
}

// NR = 1: every (KB | CB, RJ) variant (experiments); NR = 2: the tuned defaults only (compile time)
template <int MAXT, int NR>
static void launch_fwd_front_t(int kb, unsigned nf, int nt, int pstride, cudaStream_t st, int64_t first, int hc,
                               const SolveCtx& cx, const double* b) {
    const size_t sh = (size_t)NR * pstride * sizeof(double);
    if constexpr (NR == 1) {
        if (kb <= 2) launch_pdl(k_fwd_front<MAXT, 2, 1>, nf, nt, sh, st, first, hc, cx, b, pstride);
        else if (kb <= 4) launch_pdl(k_fwd_front<MAXT, 4, 1>, nf, nt, sh, st, first, hc, cx, b, pstride);
        else if (kb <= 8 || MAXT > 512) launch_pdl(k_fwd_front<MAXT, 8, 1>, nf, nt, sh, st, first, hc, cx, b, pstride);
        else launch_pdl(k_fwd_front<(MAXT > 512 ? 512 : MAXT), 16, 1>, nf, nt, sh, st, first, hc, cx, b, pstride);
    } else {
        if (kb <= 4) launch_pdl(k_fwd_front<MAXT, 4, NR>, nf, nt, sh, st, first, hc, cx, b, pstride);
        else launch_pdl(k_fwd_front<MAXT, 8, NR>, nf, nt, sh, st, first, hc, cx, b, pstride);
    }
}
template <int NR>
static void launch_fwd_front(int nt, int kb, unsigned nf, int pstride, cudaStream_t st, int64_t first, int hc,
                             const SolveCtx& cx, const double* b) {
    if (nt <= 256) launch_fwd_front_t<256, NR>(kb, nf, nt, pstride, st, first, hc, cx, b);
    else if (nt <= 512) launch_fwd_front_t<512, NR>(kb, nf, nt, pstride, st, first, hc, cx, b);
    else launch_fwd_front_t<1024, NR>(kb, nf, nt, pstride, st, first, hc, cx, b);
}

template <int MAXT, int RJ, int NR>
static void launch_bwd_front_t(int cb, unsigned nf, int nt, int mstride, cudaStream_t st, int64_t first, const SolveCtx& cx,
                               double* x) {
    const size_t sh = (size_t)NR * mstride * sizeof(double);
    if constexpr (NR == 1) {
        if (cb <= 1) launch_pdl(k_bwd_front<MAXT, 1, RJ, 1>, nf, nt, sh, st, first, cx, x, mstride);
        else if (cb <= 2) launch_pdl(k_bwd_front<MAXT, 2, RJ, 1>, nf, nt, sh, st, first, cx, x, mstride);
        else if (cb <= 4 || MAXT > 512)
            launch_pdl(k_bwd_front<MAXT, (MAXT > 512 && RJ > 2 ? 2 : 4), RJ, 1>, nf, nt, sh, st, first, cx, x, mstride);
        else launch_pdl(k_bwd_front<(MAXT > 512 ? 512 : MAXT), 8, RJ, 1>, nf, nt, sh, st, first, cx, x, mstride);
    } else {
        if (cb <= 1) launch_pdl(k_bwd_front<MAXT, 1, RJ, NR>, nf, nt, sh, st, first, cx, x, mstride);
        else launch_pdl(k_bwd_front<MAXT, 2, RJ, NR>, nf, nt, sh, st, first, cx, x, mstride);
    }
}
template <int NR>
static void launch_bwd_front(int nt, int cb, int rj, unsigned nf, int mstride, cudaStream_t st, int64_t first,
                             const SolveCtx& cx, double* x) {
    if (rj <= 2) {
        if (nt <= 256) launch_bwd_front_t<256, 2, NR>(cb, nf, nt, mstride, st, first, cx, x);
        else if (nt <= 512) launch_bwd_front_t<512, 2, NR>(cb, nf, nt, mstride, st, first, cx, x);
        else launch_bwd_front_t<1024, 2, NR>(cb, nf, nt, mstride, st, first, cx, x);
    } else {
        if (nt <= 256) launch_bwd_front_t<256, 4, NR>(cb, nf, nt, mstride, st, first, cx, x);
        else if (nt <= 512) launch_bwd_front_t<512, 4, NR>(cb, nf, nt, mstride, st, first, cx, x);
        else launch_bwd_front_t<1024, 4, NR>(cb, nf, nt, mstride, st, first, cx, x);
    }
}

constexpr int MAX_NRHS = 2;     // right-hand sides one sweep can carry

// both sweeps for NR right-hand sides stored contiguously in b / x (stride N)
template <int NR>
static void sweep_launches(ws_solver* s, const double* bi, double* xi, cudaStream_t st) {
    const Symbolic& Y = s->sym;
    static const bool flagsync_env = !(getenv("WS_NO_FLAGSYNC") && atoi(getenv("WS_NO_FLAGSYNC"))) &&
                                     !(getenv("WS_NO_PDL") && atoi(getenv("WS_NO_PDL")));
    const bool flagsync = flagsync_env && s->overlap;
    g_pdl_on = s->overlap;
    SolveCtx cx{s->recs.get(), s->S.get(), s->perm.get(), s->wsrc0.get(), s->wsrc1.get(), s->bnd.get(),
                s->dinv.get(), s->W.get(), s->yd.get(), s->x.get(), (int64_t)((Y.Wtotal + 1) & ~(int64_t)1), s->N,
                flagsync ? s->sflags.get() : nullptr, flagsync ? s->sflags.get() + Y.nfronts : nullptr};
    // (slot 0 is the sticky abort flag: front ids start at 1)
    if (flagsync) WS_CUDA(cudaMemsetAsync(s->sflags.get() + 1, 0, s->sflags.bytes() - sizeof(int32_t), st));
    const SweepTuning& tune = sweep_tuning();
    int nl = 0;
    // experiments (tools/sweep_levels.py): stop after the first WS_SWEEP_STOP launches -- the time of a truncated solve,
    // as a function of the cut, is the cost of each level INSIDE the overlapped pipeline (ncu serialises the launches)
    // (read once; WS_TUNE_RELOAD=1 re-reads it per solve, like the tuning knobs)
    static const bool stop_reload = getenv("WS_TUNE_RELOAD") && atoi(getenv("WS_TUNE_RELOAD"));
    static int stop_cached = getenv("WS_SWEEP_STOP") ? atoi(getenv("WS_SWEEP_STOP")) : 1 << 30;
    if (stop_reload) stop_cached = getenv("WS_SWEEP_STOP") ? atoi(getenv("WS_SWEEP_STOP")) : 1 << 30;
    const int stop_after = stop_cached;
    // Optional (WS_SPLIT=1): every level launched in two halves (left / right subtree of the root).  With the
    // per-front flags the left half of level l-1 depends on the left half of level l only, so it can run while the
    // right half of level l drains.  Measured: 0.977 ms against 0.970 ms per solve -- the tails are not what is
    // left to gain -- so it is off by default.
    static const bool halves_env = getenv("WS_SPLIT") && atoi(getenv("WS_SPLIT"));
    const bool halves = flagsync && halves_env;
    for (int l = Y.depth; l >= 0; --l) {
        if (nl >= stop_after) break;
        const LevelPlan& LP = s->levels[l];
        const int hc = l < Y.depth ? 1 : 0;
        const int nparts = (halves && LP.nfronts >= 2) ? 2 : 1;
        for (int part = 0; part < nparts; ++part) {
            if (LP.small) {
                const int64_t h = nparts == 2 ? LP.nfronts / 2 : LP.nfronts;
                const int64_t f = LP.first_front + (part ? h : 0);
                const int64_t n = nparts == 2 ? (part ? LP.nfronts - h : h) : LP.nfronts;
                const bool mid = LP.nfronts <= tune.mid_fronts;
                launch_fwd_front<NR>(LP.nt, l == Y.depth ? tune.kb_leaf : (mid ? tune.kb_mid : tune.kb), (unsigned)n, std::max(LP.maxp, 2), st, f, hc, cx, bi);
                ++nl;
            } else if (LP.fwd.cnt) {
                const int64_t o = LP.fwd.off + (nparts == 2 && part ? LP.fwd_mid : 0);
                const int64_t n = nparts == 2 ? (part ? LP.fwd.cnt - LP.fwd_mid : LP.fwd_mid) : LP.fwd.cnt;
                if (n <= 0) continue;
                const int base = (LP.maxp + 1) & ~1;
                if (LP.fwd.cnt >= 500 || LP.narrow_tiles)
                    launch_pdl(k_fwd_tile<8, NR>, (unsigned)n, 32 * 8, (size_t)NR * (base + 8 * 33) * sizeof(double), st,
                               s->fwd_tiles.get() + o, hc, cx, bi, base);
                else
                    launch_pdl(k_fwd_tile<32, NR>, (unsigned)n, 32 * 32, (size_t)NR * (base + 32 * 33) * sizeof(double), st,
                               s->fwd_tiles.get() + o, hc, cx, bi, base);
                ++nl;
            }
        }
    }
    for (int l = 0; l <= Y.depth; ++l) {
        if (nl >= stop_after) break;
        const LevelPlan& LP = s->levels[l];
        const int nparts = (halves && LP.nfronts >= 2) ? 2 : 1;
        for (int part = 0; part < nparts; ++part) {
            if (LP.small) {
                const int64_t h = nparts == 2 ? LP.nfronts / 2 : LP.nfronts;
                const int64_t f = LP.first_front + (part ? h : 0);
                const int64_t n = nparts == 2 ? (part ? LP.nfronts - h : h) : LP.nfronts;
                const bool mid = LP.nfronts <= tune.mid_fronts;
                launch_bwd_front<NR>(LP.nt, l == Y.depth ? tune.cb_leaf : (mid ? tune.cb_mid : tune.cb),
                                     l == Y.depth ? tune.rj_leaf : (mid ? tune.rj_mid : tune.rj), (unsigned)n, std::max(LP.maxm, 2), st, f, cx, xi);
                ++nl;
            } else if (LP.bwd.cnt) {
                const int64_t o = LP.bwd.off + (nparts == 2 && part ? LP.bwd_mid : 0);
                const int64_t n = nparts == 2 ? (part ? LP.bwd.cnt - LP.bwd_mid : LP.bwd_mid) : LP.bwd.cnt;
                if (n <= 0) continue;
                const int base = (LP.maxm + 1) & ~1;
                if (LP.bwd.cnt >= 430 || LP.narrow_tiles)
                    launch_pdl(k_bwd_tile<1, NR>, (unsigned)n, 32 * GEMVT_COLS, (size_t)NR * (base + GEMVT_COLS) * sizeof(double),
                               st, s->bwd_tiles.get() + o, cx, xi, base);
                else
                    launch_pdl(k_bwd_tile<4, NR>, (unsigned)n, 32 * GEMVT_COLS * 4,
                               (size_t)NR * (base + GEMVT_COLS * 4) * sizeof(double), st, s->bwd_tiles.get() + o, cx, xi, base);
                ++nl;
            }
        }
    }
    count_launch(nl);
}

static void configure_sweeps() {
    static PerDeviceOnce configure;
    configure([] {
        const int big = 200 * 1024;
        WS_CUDA(cudaFuncSetAttribute(k_fwd_tile<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_fwd_tile<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_bwd_tile<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_bwd_tile<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_fwd_tile<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_fwd_tile<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_bwd_tile<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        WS_CUDA(cudaFuncSetAttribute(k_bwd_tile<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    });
}

void spmv_launch(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st);

}  // namespace ws

using namespace ws;

extern "C" {

int ws_solver_create(const ws_pattern* p, const ws_symbolic* sym, ws_solver** out, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K8 solver create");
    WS_REQUIRE(out, "out == NULL");
    *out = nullptr;
    WS_REQUIRE(p && sym, "null handle");
    WS_REQUIRE(p->N == sym->S.N, "pattern and symbolic analysis have different sizes");
    DeviceGuard g(device);
    configure_mempool(device);
    cudaStream_t st = (cudaStream_t)stream;
    auto* s = new ws_solver();
    {
        static std::atomic<int64_t> next_uid{1};
        s->uid = next_uid.fetch_add(1);
    }
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
        if (sym->on_device) {
            WS_REQUIRE(sym->device == device, "symbolic analysis lives on another device");
            clone(s->perm, sym->d_perm, (size_t)Y.N, st);
            clone(s->iperm, sym->d_iperm, (size_t)Y.N, st);
            clone(s->node_of_new, sym->d_node_of_new, (size_t)Y.N, st);
            clone(s->bnd, sym->d_bnd, (size_t)sym->sumq, st);
            clone(s->rel, sym->d_rel, (size_t)sym->sumq, st);
        } else {
            upload(s->perm, Y.perm, st);
            upload(s->iperm, Y.iperm, st);
            upload(s->node_of_new, Y.node_of_new, st);
            upload(s->bnd, Y.bnd, st);
            upload(s->rel, Y.rel, st);
        }
        upload(s->fr_p, Y.p, st);
        upload(s->fr_q, Y.q, st);
        upload(s->fr_start, Y.start, st);
        upload(s->fr_Loff, Y.Loff, st);
        upload(s->fr_Uoff, Y.Uoff, st);
        upload(s->fr_woff, Y.woff, st);
        upload(s->fr_bndoff, Y.bndoff, st);
        s->L.alloc(Y.Ltotal, st);
        s->S.alloc(Y.Ltotal, st);
        s->T.alloc(Y.Ltotal, st);
        s->U.alloc(std::max<int64_t>(Y.Utotal, 2), st);
        s->W.alloc((size_t)MAX_NRHS * ((Y.Wtotal + 1) & ~(int64_t)1), st);
        s->d.alloc(Y.N, st);
        s->dinv.alloc(Y.N, st);
        s->yd.alloc((size_t)MAX_NRHS * Y.N, st);
        s->x.alloc((size_t)MAX_NRHS * Y.N, st);
        classify_levels(s);
        s->sflags.alloc((size_t)2 * Y.nfronts, st);
        WS_CUDA(cudaMemsetAsync(s->sflags.get(), 0, s->sflags.bytes(), st));
        {
            // per front: how many CTAs the forward / backward sweep spends on it (1 for the one-CTA-per-front
            // kernels, the number of row / column tiles for the large fronts): what a waiter compares with
            std::vector<FrontRec> recs((size_t)Y.nfronts);
            for (int64_t t = 1; t < Y.nfronts; ++t) {
                const int p_ = Y.p[t], m_ = p_ + Y.q[t];
                int nf = 1, nbk = 1;
                if (!s->levels[63 - __builtin_clzll((unsigned long long)t)].small) {
                    nf = 0;
                    for (int r0 = 0; r0 < m_; r0 += GEMV_ROWS) {
                        int nr = std::min(GEMV_ROWS, m_ - r0);
                        if (r0 < p_ && r0 + nr > p_) nr = p_ - r0;
                        ++nf;
                        if (nr < GEMV_ROWS && r0 + nr < m_) ++nf;
                    }
                    nbk = (p_ + GEMVT_COLS - 1) / GEMVT_COLS;
                }
                recs[t] = FrontRec{Y.Loff[t], Y.start[t], Y.woff[t], Y.bndoff[t], Y.p[t], Y.q[t], nf, nbk, 0, 0, 0, 0};
            }
            for (int64_t t = 1; t < Y.nfronts; ++t) {
                if (2 * t + 1 < Y.nfronts) {
                    recs[t].cf0 = recs[2 * t].pad0;
                    recs[t].cf1 = recs[2 * t + 1].pad0;
                }
                if (t > 1) recs[t].pb = recs[t >> 1].pad1;
            }
            upload(s->recs, recs, st);
            WS_CUDA(cudaStreamSynchronize(st));   // recs dies here
            WS_REQUIRE(Y.Wtotal < ((int64_t)1 << 31), "work-vector storage exceeds 32-bit indexing");
            s->wsrc0.alloc(std::max<int64_t>(Y.Wtotal, 1), st);
            s->wsrc1.alloc(std::max<int64_t>(Y.Wtotal, 1), st);
            WS_CUDA(cudaMemsetAsync(s->wsrc0.get(), 0xFF, s->wsrc0.bytes(), st));
            WS_CUDA(cudaMemsetAsync(s->wsrc1.get(), 0xFF, s->wsrc1.bytes(), st));
            if (Y.nfronts > 2) {
                k_build_wsrc<<<(unsigned)(Y.nfronts - 2), 128, 0, st>>>(Y.nfronts, s->recs.get(), s->rel.get(),
                                                                         s->wsrc0.get(), s->wsrc1.get());
                count_launch(1);
            }
        }
        s->stats.alloc(NSTATS, st);
        s->pat = p;
        s->dest.alloc(p->nnz, st);
        s->bytes = (3 * Y.Ltotal + Y.Utotal + 2 * Y.Wtotal + 4 * Y.N) * 8 + p->nnz * 8;
        const double tc1 = tnow();
        const int T = 128;
        k_dest_map<<<ceil_div(Y.N, T), T, 0, st>>>(Y.N, p->rowptr.get(), p->colidx.get(), s->iperm.get(),
                                                     s->node_of_new.get(), view(s), s->dest.get());
        WS_CUDA(cudaMemsetAsync(s->stats.get(), 0, NSTATS * sizeof(unsigned long long), st));
        k_count_bad_dest<<<ceil_div(p->nnz, 256), 256, 0, st>>>(p->nnz, s->dest.get(), s->stats.get());
        unsigned long long bad = 0;
        WS_CUDA(cudaMemcpyAsync(&bad, s->stats.get(), sizeof bad, cudaMemcpyDeviceToHost, st));
        count_launch(2);
        const double tc2 = tnow();
        WS_CUDA(cudaStreamSynchronize(st));
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
    int rc = ws_solver_factor_begin(s, A_vals, B_vals, sigma, device, stream);
    if (rc) return rc;
    return ws_solver_factor_end(s, device, stream);
}

int ws_solver_factor_begin(ws_solver* s, const double* A_vals, const double* B_vals, double sigma, int device,
                           void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K8b factor");
    WS_REQUIRE(s && A_vals, "null pointer");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const Symbolic& Y = s->sym;
    s->factored = false;
    configure_sweeps();
    WS_CUDA(cudaMemsetAsync(s->L.get(), 0, s->L.bytes(), st));
    WS_CUDA(cudaMemsetAsync(s->U.get(), 0, s->U.bytes(), st));
    WS_CUDA(cudaMemsetAsync(s->S.get(), 0, s->S.bytes(), st));   // blocks above the block diagonal are read as zeros
    s->refine_steps = 0;
    s->fA = A_vals;
    s->fB = B_vals;
    s->fsigma = sigma;
    unsigned long long init[NSTATS] = {0ull, 0ull, ~0ull, 0ull};
    WS_CUDA(cudaMemcpyAsync(s->stats.get(), init, sizeof init, cudaMemcpyHostToDevice, st));
    {
        int blocks = (int)std::min<int64_t>((s->nnz + 255) / 256, (int64_t)kNumSMs * 32);
        k_scatter<<<blocks, 256, 0, st>>>(s->nnz, s->dest.get(), A_vals, B_vals, sigma, s->L.get(), s->stats.get());
        count_launch(1);
    }
    FrontView fv = view(s);
    static PerDeviceOnce configure_fused;
    configure_fused([] {
        WS_CUDA(cudaFuncSetAttribute(k_front_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_MAX_SMEM));
        WS_CUDA(cudaFuncSetAttribute(k_front_fused_regs<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_MAX_SMEM));
    });
    for (int l = Y.depth; l >= 0; --l) {
        const LevelPlan& LP = s->levels[l];
        if (LP.fused) {
            FusedCtx fc{fv, s->L.get(), s->S.get(), s->U.get(), s->d.get(), s->stats.get(), s->wsrc0.get(), s->wsrc1.get()};
            const size_t sh = fused_smem_doubles(LP.maxm, LP.maxp) * sizeof(double);
            if (LP.maxp <= 48) k_front_fused<<<(unsigned)LP.nfronts, LP.nt, sh, st>>>(LP.first_front, l < Y.depth ? 1 : 0, fc);
            else k_front_fused_regs<64><<<(unsigned)LP.nfronts, LP.nt, sh, st>>>(LP.first_front, l < Y.depth ? 1 : 0, fc);
            count_launch(1);
            continue;
        }
        if (!s->plan_built) build_plan(s, st);
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
    if (!s->plan_built) build_plan(s, st);
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
    // ---- accuracy guard -------------------------------------------------------------------------------
    // The factorisation does not pivot.  That is backward stable for the definite scalar matrix at the default
    // shift and, in practice, for the quasi-definite vector matrix; with target_neff the shift sits inside the
    // spectrum (fe_solver.py:313-316; getting-started.ipynb cell 26 uses target_neff = 0.995 under every index
    // of the fibre) and A - sigma B is genuinely indefinite.  So every factorisation is CHECKED instead of
    // trusted: tiny pivots are replaced by +-tau (above), then a probe system M x = b is solved with the new
    // factors and its normwise backward error eta = |b - M x| / (max|M| |x| + |b|) decides how many steps of
    // iterative refinement every later solve performs (0 when the factors pass as they are -- the default-shift
    // cost of the guard is this one solve + one SpMV).  If refinement cannot bring eta below ETA_FAIL the call
    // fails with WS_ERR_NUMERIC: degraded eigenpairs are never returned silently.
    static const bool guard = !(getenv("WS_NO_GUARD") && atoi(getenv("WS_NO_GUARD")));
    s->factored = true;                       // the probe goes through the regular sweeps
    const int T = 256;
    const int nbN = ceil_div(s->N, T);
    auto probe_pass = [&](bool first) {
        // residual of the current iterate rf_x against rf_b, norms -> stats[6..8]
        spmv_launch(s->pat, s->fA, s->rf_x.get(), s->rf_y.get(), st);
        const double* yb = nullptr;
        if (s->fB) {
            spmv_launch(s->pat, s->fB, s->rf_x.get(), s->rf_dx.get(), st);
            yb = s->rf_dx.get();
        }
        if (!first) WS_CUDA(cudaMemsetAsync(s->stats.get() + 6, 0, 3 * sizeof(unsigned long long), st));
        k_residual<<<nbN, T, 0, st>>>(s->N, s->rf_b.get(), s->rf_y.get(), yb, s->fsigma, s->rf_x.get(), s->rf_r.get(),
                                      s->stats.get() + 6);
        count_launch(1);
    };
    if (guard && s->N > 0) {
        if (!s->rf_b.get()) {
            s->rf_b.alloc(s->N, st); s->rf_r.alloc(s->N, st); s->rf_y.alloc(s->N, st); s->rf_dx.alloc(s->N, st); s->rf_x.alloc(s->N, st);
            s->bytes += 5 * s->N * 8;
        }
        k_probe_rhs<<<nbN, T, 0, st>>>(s->N, s->rf_b.get());
        count_launch(1);
        sweep_launches<1>(s, s->rf_b.get(), s->rf_x.get(), st);
        probe_pass(true);
    }
    if (!s->h_stats) WS_CUDA(cudaMallocHost((void**)&s->h_stats, NSTATS * sizeof(unsigned long long)));
    WS_CUDA(cudaMemcpyAsync(s->h_stats, s->stats.get(), NSTATS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    s->factored = false;
    s->factor_pending = true;
    WS_API_END
}

int ws_solver_factor_end(ws_solver* s, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K8b factor: statistics + accuracy probe");
    WS_REQUIRE(s, "solver == NULL");
    if (!s->factor_pending) throw Error(WS_ERR_STATE, "ws_solver_factor_end without ws_solver_factor_begin");
    s->factor_pending = false;
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool guard = !(getenv("WS_NO_GUARD") && atoi(getenv("WS_NO_GUARD")));
    // Refinement starts at a probe backward error of 1e-14 (it was 1e-13): the symmetric iteration (kind 0) needs the solve to
    // be a symmetric operator to well below its convergence tolerance, and an unrefined solve at eta ~ 1e-13 is not -- the
    // hollow-core fibre at target_neff = 0.995 (tests/test_gpu_robust.py) sat at eta0 = 1.05e-13, refined and converged in 13
    // restarts, but any change of rounding in the pivot blocks that moved eta0 to 0.9e-13 left it unrefined and stagnating
    // (2 of 6 pairs in 300 restarts).  Default shifts sit at 1e-17 .. 2e-15 and are not affected.
    static const double ETA_OK = getenv("WS_ETA_OK") ? atof(getenv("WS_ETA_OK")) : 1e-14;
    constexpr double ETA_FAIL = 1e-11;
    constexpr int MAX_REFINE = 5;
    const int T = 256;
    const int nbN = ceil_div(s->N, T);
    auto probe_pass = [&](bool first) {
        spmv_launch(s->pat, s->fA, s->rf_x.get(), s->rf_y.get(), st);
        const double* yb = nullptr;
        if (s->fB) {
            spmv_launch(s->pat, s->fB, s->rf_x.get(), s->rf_dx.get(), st);
            yb = s->rf_dx.get();
        }
        if (!first) WS_CUDA(cudaMemsetAsync(s->stats.get() + 6, 0, 3 * sizeof(unsigned long long), st));
        k_residual<<<nbN, T, 0, st>>>(s->N, s->rf_b.get(), s->rf_y.get(), yb, s->fsigma, s->rf_x.get(), s->rf_r.get(),
                                      s->stats.get() + 6);
        count_launch(1);
    };
    unsigned long long* h = s->h_stats;
    WS_CUDA(cudaStreamSynchronize(st));
    WS_CUDA(cudaGetLastError());
    auto as_double = [](unsigned long long bits) { double d; std::memcpy(&d, &bits, 8); return d; };
    s->h_neg = (int64_t)h[0];
    s->h_bad = (int64_t)h[1];
    s->h_minpiv = h[2] == ~0ull ? 0.0 : as_double(h[2]);
    s->h_maxpiv = as_double(h[3]);
    s->h_pert = (int64_t)h[4];
    s->h_maxabs = as_double(h[5]);
    s->probe_iters = 0;
    s->h_eta0 = s->h_eta = 0.0;
    if (s->h_bad)
        throw Error(WS_ERR_NUMERIC, "factorisation met " + std::to_string(s->h_bad) + " non-finite pivots (matrix entries not finite?)");
    if (guard && s->N > 0) {
        auto eta_of = [&](const unsigned long long* hh) {
            const double r = as_double(hh[6]), x = as_double(hh[7]), bb = as_double(hh[8]);
            const double den = s->h_maxabs * x + bb;
            return den > 0.0 && std::isfinite(r) && std::isfinite(x) ? r / den : INFINITY;
        };
        double eta = eta_of(h);
        s->h_eta0 = eta;
        int steps = 0;
        while (eta > ETA_OK && steps < MAX_REFINE) {
            // one refinement step on the probe: x += M~^-1 (b - M x)
            s->factored = true;
            sweep_launches<1>(s, s->rf_r.get(), s->rf_y.get(), st);
            s->factored = false;
            k_add_inplace<<<nbN, T, 0, st>>>(s->N, s->rf_x.get(), s->rf_y.get());
            count_launch(1);
            probe_pass(false);
            WS_CUDA(cudaMemcpyAsync(h, s->stats.get(), NSTATS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            WS_CUDA(cudaStreamSynchronize(st));
            const double e2 = eta_of(h);
            ++steps;
            const bool stalled = !(e2 < 0.25 * eta);
            eta = std::min(eta, e2);
            if (stalled) break;
        }
        s->probe_iters = steps;
        s->h_eta = eta;
        if (getenv("WS_TRACE"))
            fprintf(stderr, "[ws_solver_factor] guard: perturbed pivots %lld, max|M| %.3e, eta0 %.2e -> eta %.2e after %d refinement steps\n",
                    (long long)s->h_pert, s->h_maxabs, s->h_eta0, eta, steps);
        if (!(eta <= ETA_FAIL))
            throw Error(WS_ERR_NUMERIC, "factorisation of A - sigma B is not accurate at this shift: probe backward error " +
                                            std::to_string(eta) + " after " + std::to_string(steps) + " refinement steps, " +
                                            std::to_string(s->h_pert) + " perturbed pivots (shift too close to an eigenvalue, or the "
                                            "pivot-free LDL^T is unstable for this matrix)");
        // one spare step when refinement was needed at all: other right-hand sides may converge a little slower
        s->refine_steps = steps ? std::min(steps + 1, MAX_REFINE) : 0;
    }
    s->factored = true;
    WS_API_END
}

// one refined solve: x = M~^-1 b, then `steps` times x += M~^-1 (b - M x) with M applied as SpMV on the pattern
static void refined_solve(ws_solver* s, const double* b, double* x, cudaStream_t st) {
    const int T = 256;
    const int nbN = ceil_div(s->N, T);
    sweep_launches<1>(s, b, x, st);
    for (int it = 0; it < s->refine_steps; ++it) {
        spmv_launch(s->pat, s->fA, x, s->rf_y.get(), st);
        const double* yb = nullptr;
        if (s->fB) {
            spmv_launch(s->pat, s->fB, x, s->rf_dx.get(), st);
            yb = s->rf_dx.get();
        }
        k_residual<<<nbN, T, 0, st>>>(s->N, b, s->rf_y.get(), yb, s->fsigma, x, s->rf_r.get(), nullptr);
        sweep_launches<1>(s, s->rf_r.get(), s->rf_y.get(), st);
        k_add_inplace<<<nbN, T, 0, st>>>(s->N, x, s->rf_y.get());
        count_launch(2);
    }
}

int ws_solver_solve(ws_solver* s, const double* b, double* x, int nrhs, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K8c solve sweeps");
    WS_REQUIRE(s && b && x && nrhs >= 1, "bad arguments");
    if (!s->factored) throw Error(WS_ERR_STATE, "ws_solver_solve before a successful ws_solver_factor");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    // (b may alias x: b is read by the forward sweep only, x is written by the backward sweep only)
    configure_sweeps();
    if (s->refine_steps > 0) {
        // (refinement reads b again after x has been written: aliasing is resolved through a copy)
        for (int rhs = 0; rhs < nrhs; ++rhs) {
            const double* bi = b + (size_t)rhs * s->N;
            double* xi = x + (size_t)rhs * s->N;
            if (bi == xi) {
                WS_CUDA(cudaMemcpyAsync(s->rf_b.get(), bi, s->N * sizeof(double), cudaMemcpyDeviceToDevice, st));
                bi = s->rf_b.get();
            }
            refined_solve(s, bi, xi, st);
        }
        WS_CUDA(cudaGetLastError());
        return WS_OK;
    }
    int rhs = 0;
    for (; rhs + 2 <= nrhs; rhs += 2) sweep_launches<2>(s, b + (size_t)rhs * s->N, x + (size_t)rhs * s->N, st);
    for (; rhs < nrhs; ++rhs) sweep_launches<1>(s, b + (size_t)rhs * s->N, x + (size_t)rhs * s->N, st);
    WS_CUDA(cudaGetLastError());
    WS_API_END
}

// Host check of the abort flag of the flag-synchronised sweeps (set when a wait ran into its bound:
// a logic error, never expected).  Synchronises the stream.
int ws_solver_check(const ws_solver* s, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(s, "solver == NULL");
    DeviceGuard g(device);
    int32_t flag = 0;
    if (s->sflags.get()) {
        WS_CUDA(cudaMemcpyAsync(&flag, s->sflags.get(), sizeof flag, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        WS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    if (flag) {
        WS_CUDA(cudaMemsetAsync(s->sflags.get(), 0, sizeof(int32_t), (cudaStream_t)stream));
        throw Error(WS_ERR_STATE, "internal: a solve sweep timed out waiting for a front (results are invalid)");
    }
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

int ws_solver_set_overlap(ws_solver* s, int on) {
    WS_API_BEGIN
    WS_REQUIRE(s, "solver == NULL");
    s->overlap = on != 0;
    WS_API_END
}

int ws_solver_quality(const ws_solver* s, int64_t* h_info, double* h_dinfo) {
    WS_API_BEGIN
    WS_REQUIRE(s, "solver == NULL");
    if (h_info) {
        h_info[0] = s->h_pert; h_info[1] = s->refine_steps; h_info[2] = s->probe_iters; h_info[3] = s->uid;
    }
    if (h_dinfo) {
        h_dinfo[0] = s->h_eta0; h_dinfo[1] = s->h_eta; h_dinfo[2] = s->h_maxabs; h_dinfo[3] = PIV_REL * s->h_maxabs;
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
