// K5: CSR sparse matrix-vector product, fp64 (the B*x and (A - sigma B)*x applications that
// ARPACK requests through reverse communication inside eigsh, fe_solver.py:328).
//
// FEM rows are short (P2 scalar: 9 or ~19 entries; vector: 9 or ~19) so a sub-warp of LANES
// threads owns a row: the lanes read consecutive (value, column) pairs -- coalesced 8 B + 4 B
// streams, the only compulsory HBM traffic -- gather x through L2 (x is N*8 B, L2-resident for
// every mesh in scope) and reduce with shuffles.  Bound: HBM, 12 B per non-zero + 20 B per row.
#include "ws_common.cuh"

namespace ws {

template <int LANES>
__global__ void __launch_bounds__(256)
k_spmv_csr(int64_t N, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
           const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t row = gt / LANES;
    const int lane = (int)(gt % LANES);
    double acc = 0.0;
    if (row < N) {
        const int s1 = rowptr[row + 1];
        for (int s = rowptr[row] + lane; s < s1; s += LANES)
            acc = fma(__ldg(vals + s), __ldg(x + __ldg(colidx + s)), acc);
    }
#pragma unroll
    for (int o = LANES / 2; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LANES);
    if (row < N && lane == 0) y[row] = acc;
}

void spmv_launch(const ws_pattern* p, const double* vals, const double* x, double* y, cudaStream_t st) {
    if (p->N == 0) return;
    const double avg = p->N ? (double)p->nnz / (double)p->N : 0.0;
    const int T = 256;
    if (avg > 24) {
        k_spmv_csr<16><<<ceil_div(p->N * 16, T), T, 0, st>>>(p->N, p->rowptr.get(), p->colidx.get(), vals, x, y);
    } else if (avg > 6) {
        k_spmv_csr<8><<<ceil_div(p->N * 8, T), T, 0, st>>>(p->N, p->rowptr.get(), p->colidx.get(), vals, x, y);
    } else {
        k_spmv_csr<4><<<ceil_div(p->N * 4, T), T, 0, st>>>(p->N, p->rowptr.get(), p->colidx.get(), vals, x, y);
    }
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
    DeviceGuard g(device);
    spmv_launch(p, vals, x, y, (cudaStream_t)stream);
    WS_API_END
}
