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
#include "ws_common.cuh"

namespace ws {

constexpr int SPMV_ROWS = 128;          // rows (and threads) per CTA
constexpr int SPMV_CAP = 4096;          // products held in shared memory (32 KB)

__global__ void __launch_bounds__(SPMV_ROWS)
k_spmv_stream(int64_t N, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
              const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
    __shared__ double prod[SPMV_CAP];
    __shared__ int rp[SPMV_ROWS + 1];
    const int64_t r0 = (int64_t)blockIdx.x * SPMV_ROWS;
    const int nr = (int)min((int64_t)SPMV_ROWS, N - r0);
    const int tid = threadIdx.x;
    if (tid < nr) rp[tid] = rowptr[r0 + tid];
    if (tid == 0) rp[nr] = rowptr[r0 + nr];
    __syncthreads();
    const int p0 = rp[0], p1 = rp[nr];
    const int cnt = p1 - p0;
    if (cnt <= SPMV_CAP) {
        int s = tid;
        for (; s + 3 * SPMV_ROWS < cnt; s += 4 * SPMV_ROWS) {
            const int c0 = __ldg(colidx + p0 + s), c1 = __ldg(colidx + p0 + s + SPMV_ROWS);
            const int c2 = __ldg(colidx + p0 + s + 2 * SPMV_ROWS), c3 = __ldg(colidx + p0 + s + 3 * SPMV_ROWS);
            const double v0 = __ldg(vals + p0 + s), v1 = __ldg(vals + p0 + s + SPMV_ROWS);
            const double v2 = __ldg(vals + p0 + s + 2 * SPMV_ROWS), v3 = __ldg(vals + p0 + s + 3 * SPMV_ROWS);
            prod[s] = v0 * __ldg(x + c0);
            prod[s + SPMV_ROWS] = v1 * __ldg(x + c1);
            prod[s + 2 * SPMV_ROWS] = v2 * __ldg(x + c2);
            prod[s + 3 * SPMV_ROWS] = v3 * __ldg(x + c3);
        }
        for (; s < cnt; s += SPMV_ROWS) prod[s] = __ldg(vals + p0 + s) * __ldg(x + __ldg(colidx + p0 + s));
        __syncthreads();
        if (tid < nr) {
            double acc = 0.0;
            const int a = rp[tid] - p0, b = rp[tid + 1] - p0;
            for (int i = a; i < b; ++i) acc += prod[i];
            y[r0 + tid] = acc;
        }
    } else if (tid < nr) {      // unusually dense block: plain row loop
        double acc = 0.0;
        for (int i = rp[tid]; i < rp[tid + 1]; ++i) acc += vals[i] * x[colidx[i]];
        y[r0 + tid] = acc;
    }
}

void spmv_launch(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st) {
    if (p->N == 0) return;
    k_spmv_stream<<<ceil_div(p->N, SPMV_ROWS), SPMV_ROWS, 0, st>>>(p->N, p->rowptr.get(), p->colidx.get(), vals, x, y);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
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
