// K5: CSR sparse matrix-vector product, fp64 (the B*x and (A - sigma B)*x applications that
// ARPACK requests through reverse communication inside eigsh, fe_solver.py:328).
//
// FEM rows are short (9 or ~19 entries), so a "one sub-warp per row" kernel is latency bound:
// every lane has only one or two loads in flight.  This kernel streams instead: a CTA owns a
// block of consecutive rows, i.e. one contiguous slice of (value, column) pairs; all threads
// walk that slice with unit stride (fully coalesced 8 B + 4 B streams -- the only compulsory
// HBM traffic -- and ~12 independent gathers of x per thread in flight), park the products in
// shared memory, and one thread per row then adds its products in column order (deterministic).
// x is N*8 B and stays L2-resident for every mesh in scope.
// Bound: HBM, 12 B per non-zero + 20 B per row.
#include <algorithm>

#include "ws_common.cuh"

namespace ws {

constexpr int SPMV_ROWS = 128;          // rows (and threads) per CTA
constexpr int SPMV_CHUNKS = 3;          // 4-entry chunks a thread keeps in flight per round (12 entries)
constexpr int SPMV_MAX_SMEM = 64 * 1024;

// One round = every thread issues SPMV_CHUNKS 16-byte column loads and 2*SPMV_CHUNKS 16-byte value
// loads back to back (the slice is walked from its 4-entry-aligned start, so all of them are
// aligned vector loads and a warp covers 512 + 1024 contiguous bytes per chunk), then the 12
// gathers of x, then parks the products: the dependent chain of a CTA is
// rowptr -> (col, val) -> x -> shared memory, once for a block of edge / mid-side rows (~9
// entries per row) and twice for a block of vertex / node rows (~19 per row).
// NV vectors (x with stride ldx, y with stride ldy) share the (value, column) stream: a second vector costs its own
// gathers and products only.
template <int NV>
__global__ void __launch_bounds__(SPMV_ROWS)
k_spmv_stream(int64_t N, int64_t nnz, int cap, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
              const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx, double* __restrict__ y, int64_t ldy) {
    extern __shared__ double prod[];          // [NV][cap]
    __shared__ int rp[SPMV_ROWS + 1];
    const int64_t r0 = (int64_t)blockIdx.x * SPMV_ROWS;
    const int nr = (int)min((int64_t)SPMV_ROWS, N - r0);
    const int tid = threadIdx.x;
    if (tid < nr) rp[tid] = rowptr[r0 + tid];
    if (tid == 0) rp[nr] = rowptr[r0 + nr];
    __syncthreads();
    const int p0 = rp[0], p1 = rp[nr];
    const int cnt = p1 - p0;
    if (cnt <= cap) {
        const int a0 = p0 & ~3;                       // aligned start of the slice
        const int nchunk = (p1 - a0 + 3) >> 2;
        for (int cb = 0; cb < nchunk; cb += SPMV_ROWS * SPMV_CHUNKS) {
            int4 c[SPMV_CHUNKS];
            double2 va[SPMV_CHUNKS], vb[SPMV_CHUNKS];
#pragma unroll
            for (int i = 0; i < SPMV_CHUNKS; ++i) {
                const int ch = cb + tid + i * SPMV_ROWS;
                const int64_t g = (int64_t)a0 + 4 * (int64_t)ch;       // global entry index of the chunk
                c[i] = make_int4(0, 0, 0, 0);
                va[i] = vb[i] = make_double2(0.0, 0.0);
                if (ch < nchunk) {
                    if (g + 4 <= nnz) {
                        c[i] = __ldg(reinterpret_cast<const int4*>(colidx + g));
                        va[i] = __ldg(reinterpret_cast<const double2*>(vals + g));
                        vb[i] = __ldg(reinterpret_cast<const double2*>(vals + g + 2));
                    } else {                                           // last chunk of the matrix
                        int cc[4] = {0, 0, 0, 0};
                        double vv[4] = {0.0, 0.0, 0.0, 0.0};
                        for (int e = 0; e < 4; ++e)
                            if (g + e < nnz) { cc[e] = colidx[g + e]; vv[e] = vals[g + e]; }
                        c[i] = make_int4(cc[0], cc[1], cc[2], cc[3]);
                        va[i] = make_double2(vv[0], vv[1]);
                        vb[i] = make_double2(vv[2], vv[3]);
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double* xv = x + (size_t)v * ldx;
                double* pv = prod + (size_t)v * cap;
#pragma unroll
                for (int i = 0; i < SPMV_CHUNKS; ++i) {
                    // entries outside [p0, p1) belong to neighbouring blocks: their columns are valid
                    // indices, the products are simply not stored
                    const double q0 = va[i].x * __ldg(xv + c[i].x), q1 = va[i].y * __ldg(xv + c[i].y);
                    const double q2 = vb[i].x * __ldg(xv + c[i].z), q3 = vb[i].y * __ldg(xv + c[i].w);
                    const int s = (a0 - p0) + 4 * (cb + tid + i * SPMV_ROWS);   // slot of the chunk's first entry
                    if (s >= 0 && s + 3 < cnt) {
                        pv[s] = q0; pv[s + 1] = q1; pv[s + 2] = q2; pv[s + 3] = q3;
                    } else {
                        if (s >= 0 && s < cnt) pv[s] = q0;
                        if (s + 1 >= 0 && s + 1 < cnt) pv[s + 1] = q1;
                        if (s + 2 >= 0 && s + 2 < cnt) pv[s + 2] = q2;
                        if (s + 3 >= 0 && s + 3 < cnt) pv[s + 3] = q3;
                    }
                }
            }
        }
        __syncthreads();
        if (tid < nr) {
            const int a = rp[tid] - p0, b = rp[tid + 1] - p0;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double* pv = prod + (size_t)v * cap;
                double acc = 0.0;
                for (int i = a; i < b; ++i) acc += pv[i];
                y[r0 + tid + (size_t)v * ldy] = acc;
            }
        }
    } else if (tid < nr) {      // unusually dense block: plain row loop
        for (int v = 0; v < NV; ++v) {
            double acc = 0.0;
            for (int i = rp[tid]; i < rp[tid + 1]; ++i) acc += vals[i] * x[colidx[i] + (size_t)v * ldx];
            y[r0 + tid + (size_t)v * ldy] = acc;
        }
    }
}

template <int NV>
static void spmv_launch_nv(const ws_pattern* p, const double* vals, const double* x, int64_t ldx, double* y, int64_t ldy, cudaStream_t st) {
    if (p->N == 0) return;
    static PerDeviceOnce configure;
    configure([] { WS_CUDA(cudaFuncSetAttribute(k_spmv_stream<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPMV_MAX_SMEM)); });
    // shared-memory tile sized for this pattern's densest 128-row block (known from K2)
    int cap = std::min(std::max(p->max_block_slots, 256), SPMV_MAX_SMEM / 8 / NV);
    cap = (cap + 63) & ~63;
    if ((size_t)cap * NV * sizeof(double) > (size_t)SPMV_MAX_SMEM) cap = (SPMV_MAX_SMEM / 8 / NV) & ~63;
    k_spmv_stream<NV><<<ceil_div(p->N, SPMV_ROWS), SPMV_ROWS, (size_t)cap * NV * sizeof(double), st>>>(
        p->N, p->nnz, cap, p->rowptr.get(), p->colidx.get(), vals, x, ldx, y, ldy);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
}

void spmv_launch(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st) {
    spmv_launch_nv<1>(p, vals, x, p->N, y, p->N, st);
}

// two vectors stored with stride N
void spmv_launch2(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st) {
    spmv_launch_nv<2>(p, vals, x, p->N, y, p->N, st);
}

// nv = 1 or 2 vectors with independent strides (rows of a mode matrix in, scratch columns out)
void spmv_launch_strided(const ws_pattern* p, const double* vals, const double* x, int64_t ldx, double* y, int64_t ldy,
                         int nv, cudaStream_t st) {
    if (nv == 2) spmv_launch_nv<2>(p, vals, x, ldx, y, ldy, st);
    else spmv_launch_nv<1>(p, vals, x, ldx, y, ldy, st);
}

}  // namespace ws

using namespace ws;

extern "C" int ws_spmv(const ws_pattern* p, const double* vals, const double* x, double* y, int device,
                       void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    if (p->N == 0) return WS_OK;
    WS_REQUIRE(x && y && (vals || p->nnz == 0), "null pointer");
    WS_REQUIRE(x != y, "in-place product is not supported");
    DeviceGuard g(device);
    spmv_launch(p, vals, x, y, (cudaStream_t)stream);
    WS_API_END
}
