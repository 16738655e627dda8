// K16/K17: overlaps of modes in the mass-matrix inner product (second widening step).
//
// The reference hands the user `construct_B` (fe_solver.py:71-101) "for inner products" and leaves
// `u.T @ B @ v` (notes eq. 38) to scipy on the host, one pair at a time.  Here the modes are already in
// HBM (rows of the eigenvector matrix), so a whole overlap matrix G = U B V^T is two kernels per PAIR of
// columns:
//   K5  Y[0:2] = B V_j, B V_{j+1}     -- the streaming SpMV, (value, column) stream read once for both
//   K16 partial[c][i][r] = sum over chunk c (256 unknowns, one warp) of U_i[n] * Y_r[n]   (all nu rows of U
//       against both Y, which sit in registers)
// and one deterministic reduction at the end (K17, fixed lane / shuffle order: same bits on every run).
// Bound: HBM -- per pair of columns 12 B per non-zero + (nu + 4) * 8 B per unknown.
#include <algorithm>

#include "ws_common.cuh"

namespace ws {

void spmv_launch_strided(const ws_pattern* p, const double* vals, const double* x, int64_t ldx, double* y, int64_t ldy,
                         int nv, cudaStream_t st);

constexpr int OV_CHUNK = 256;      // unknowns per WARP: a lane keeps 8 entries of each Y vector in registers

// partial[((c * nu + i) * 2 + r)] = sum_{n in chunk c} U_i[n] * Y_r[n].  One warp per chunk of 256 unknowns:
// the chunk's Y entries are loaded once into registers, then every row of U streams through them with
// eight independent coalesced loads per lane (work per warp does not depend on nu mod anything).
template <int NY>
__global__ void __launch_bounds__(256) k_gram_partial(const double* __restrict__ U, int64_t ldu, int nu,
                                                      const double* __restrict__ Y, int64_t ldy, int64_t N,
                                                      int nchunks, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= nchunks) return;
    const int64_t n0 = c * OV_CHUNK + lane;
    double yv[NY][8];
#pragma unroll
    for (int r = 0; r < NY; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) yv[r][k] = (n0 + 32 * k < N) ? __ldg(Y + (size_t)r * ldy + n0 + 32 * k) : 0.0;
    const bool full = c * OV_CHUNK + OV_CHUNK <= N;
#pragma unroll 2
    for (int v = 0; v < nu; ++v) {
        const double* u = U + (size_t)v * ldu + n0;
        double uv[8];
        if (full) {
#pragma unroll
            for (int k = 0; k < 8; ++k) uv[k] = __ldg(u + 32 * k);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) uv[k] = (n0 + 32 * k < N) ? __ldg(u + 32 * k) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < NY; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += uv[k] * yv[r][k];
#pragma unroll
            for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) partial[((size_t)c * nu + v) * 2 + r] = acc;
        }
    }
}

// G[i * ldg + j0 + r] = sum_c partial[c][i][r]; one warp per (i, r)
__global__ void __launch_bounds__(32) k_gram_reduce(const double* __restrict__ partial, int nchunks, int nu, int ny,
                                                    double* __restrict__ G, int ldg, int j0) {
    const int i = blockIdx.x / ny, r = blockIdx.x % ny, lane = threadIdx.x;
    double s = 0.0;
    for (int c = lane; c < nchunks; c += 32) s += partial[((size_t)c * nu + i) * 2 + r];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) G[(size_t)i * ldg + j0 + r] = s;
}

// d[i] = partial-sum diagonal: row i0 + r of U against Y_r only (B-norms)
__global__ void __launch_bounds__(32) k_gram_reduce_diag(const double* __restrict__ partial, int nchunks, int ny,
                                                         double* __restrict__ d, int i0) {
    const int r = blockIdx.x, lane = threadIdx.x;
    double s = 0.0;
    for (int c = lane; c < nchunks; c += 32) s += partial[((size_t)c * ny + r) * 2 + r];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) d[i0 + r] = s;
}

// U_i *= 1 / sqrt(d[i]) (IEEE sqrt and division, like numpy's u / np.sqrt(u @ B @ u)); rows with d <= 0 or
// NaN are left untouched (the caller reads d and decides)
__global__ void __launch_bounds__(256) k_scale_modes(double* __restrict__ U, int64_t ldu, int64_t N,
                                                     const double* __restrict__ d) {
    const int i = blockIdx.y;
    const double di = d[i];
    if (!(di > 0.0)) return;
    const double nrm = sqrt(di);
    double* u = U + (size_t)i * ldu;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x)
        u[n] = u[n] / nrm;
}

static void gram_partial(const double* U, int64_t ldu, int nu, const double* Y, int64_t ldy, int ny, int64_t N,
                         double* partial, int nchunks, cudaStream_t st) {
    const int ctas = ceil_div(nchunks, 8);
    if (ny == 2) k_gram_partial<2><<<ctas, 256, 0, st>>>(U, ldu, nu, Y, ldy, N, nchunks, partial);
    else k_gram_partial<1><<<ctas, 256, 0, st>>>(U, ldu, nu, Y, ldy, N, nchunks, partial);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
}

}  // namespace ws

using namespace ws;

extern "C" int ws_overlap(const ws_pattern* p, const double* vals, const double* U, int64_t ldu, int nu,
                          const double* V, int64_t ldv, int nv, double* G, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    WS_REQUIRE(nu >= 0 && nv >= 0, "negative mode count");
    if (nu == 0 || nv == 0) return WS_OK;
    WS_REQUIRE(U && V && G && (vals || p->nnz == 0), "null pointer");
    WS_REQUIRE(ldu >= p->N && ldv >= p->N, "row stride smaller than the number of unknowns");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = p->N;
    if (N == 0) {
        WS_CUDA(cudaMemsetAsync(G, 0, (size_t)nu * nv * sizeof(double), st));
        return WS_OK;
    }
    configure_mempool(device);
    const int nchunks = ceil_div(N, OV_CHUNK);
    DevBuf<double> Y((size_t)2 * N, st), partial((size_t)nchunks * nu * 2, st);
    for (int j = 0; j < nv; j += 2) {
        const int ny = std::min(2, nv - j);
        spmv_launch_strided(p, vals, V + (size_t)j * ldv, ldv, Y.get(), N, ny, st);
        gram_partial(U, ldu, nu, Y.get(), N, ny, N, partial.get(), nchunks, st);
        k_gram_reduce<<<nu * ny, 32, 0, st>>>(partial.get(), nchunks, nu, ny, G, nv, j);
        WS_CUDA(cudaGetLastError());
        count_launch(1);
    }
    WS_API_END
}

extern "C" int ws_bnormalize(const ws_pattern* p, const double* vals, double* U, int64_t ldu, int nu, double* d,
                             int scale, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    WS_REQUIRE(nu >= 0, "negative mode count");
    if (nu == 0) return WS_OK;
    WS_REQUIRE(U && d && (vals || p->nnz == 0), "null pointer");
    WS_REQUIRE(ldu >= p->N, "row stride smaller than the number of unknowns");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = p->N;
    if (N == 0) {
        WS_CUDA(cudaMemsetAsync(d, 0, (size_t)nu * sizeof(double), st));
        return WS_OK;
    }
    configure_mempool(device);
    const int nchunks = ceil_div(N, OV_CHUNK);
    DevBuf<double> Y((size_t)2 * N, st), partial((size_t)nchunks * 2 * 2, st);
    for (int i = 0; i < nu; i += 2) {
        const int ny = std::min(2, nu - i);
        spmv_launch_strided(p, vals, U + (size_t)i * ldu, ldu, Y.get(), N, ny, st);
        gram_partial(U + (size_t)i * ldu, ldu, ny, Y.get(), N, ny, N, partial.get(), nchunks, st);
        k_gram_reduce_diag<<<ny, 32, 0, st>>>(partial.get(), nchunks, ny, d, i);
        WS_CUDA(cudaGetLastError());
        count_launch(1);
    }
    if (scale) {
        const int bx = std::min(ceil_div(N, 256), 4 * kNumSMs);
        k_scale_modes<<<dim3(bx, nu), 256, 0, st>>>(U, ldu, N, d);
        WS_CUDA(cudaGetLastError());
        count_launch(1);
    }
    WS_API_END
}
