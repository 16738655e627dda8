// K13-K15: resampling of modes from the triangular mesh onto points (SURVEY.md section 8(f),
// rank 1: the immediate consumer of the eigenvectors).
//
//   K13  point location.  The reference asks a KD-tree of the triangle centroids for the 1st,
//        2nd, ... nearest centroid and takes the first triangle whose inclusive inside-test holds
//        (fe_solver.py:848-863, one tree query PER TRY), i.e. the containing triangle with the
//        nearest centroid; its brute-force variant takes the first containing triangle in mesh
//        order (fe_solver.py:615-633).  Here: a uniform bin grid over the mesh (built once per
//        mesh: count -> scan -> fill), one thread per point walks the triangles registered in the
//        point's bin, applies the SAME inside-test arithmetic (this file is compiled with
//        -fmad=false, every operation rounds separately like the Python floats of
//        fe_solver.py:594-612) and keeps the best candidate under either rule.
//   K14  interpolation weights: affine map to the reference triangle (shape_funcs.py:36-49), P2
//        (shape_funcs.py:5-10), P1 (shape_funcs.py:286-288) or Whitney edge functions
//        (shape_funcs.py:331-353, signs from shape_funcs.py:317-322).
//   K15  gather + weighted sum (fe_solver.py:741-744, 770-773), real or complex fields.
//
// Bound: these are gather kernels over a few hundred bytes per point; the reference spends
// milliseconds of Python per point.
#include <algorithm>
#include <cub/cub.cuh>

#include "ws_common.cuh"

struct ws_locator {
    int device = 0;
    int64_t Ne = 0;
    int nx = 1, ny = 1;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0, inv_hx = 0, inv_hy = 0;
    ws::DevBuf<int32_t> cell_ptr;    // [nx*ny + 1]
    ws::DevBuf<int32_t> cell_tri;    // triangles registered in each bin
};

namespace ws {

namespace {

struct Grid {
    int nx, ny;
    double x0, y0, inv_hx, inv_hy;
};

__device__ __forceinline__ int cell_x(const Grid& g, double x) {
    const double t = (x - g.x0) * g.inv_hx;
    int c = t > 0.0 ? (t < (double)g.nx ? (int)t : g.nx - 1) : 0;
    return c;
}
__device__ __forceinline__ int cell_y(const Grid& g, double y) {
    const double t = (y - g.y0) * g.inv_hy;
    int c = t > 0.0 ? (t < (double)g.ny ? (int)t : g.ny - 1) : 0;
    return c;
}

// mode 0: count, mode 1: fill
__global__ void k_bin_tris(int mode, int64_t Ne, const int32_t* __restrict__ conn, int stride, const double2* __restrict__ xy,
                           Grid g, int32_t* __restrict__ count, const int32_t* __restrict__ ptr, int32_t* __restrict__ out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= Ne) return;
    const int32_t* t = conn + e * stride;
    const double2 a = xy[t[0]], b = xy[t[1]], c = xy[t[2]];
    const int cx0 = cell_x(g, fmin(a.x, fmin(b.x, c.x))), cx1 = cell_x(g, fmax(a.x, fmax(b.x, c.x)));
    const int cy0 = cell_y(g, fmin(a.y, fmin(b.y, c.y))), cy1 = cell_y(g, fmax(a.y, fmax(b.y, c.y)));
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            const int cell = cy * g.nx + cx;
            const int k = atomicAdd(&count[cell], 1);
            if (mode == 1) out[ptr[cell] + k] = (int32_t)e;
        }
}

__device__ __forceinline__ double det2(double u0, double u1, double v0, double v1) { return u0 * v1 - u1 * v0; }

// fe_solver.py:597-612 with include_vertex=True
__device__ __forceinline__ bool isinside(double px, double py, double2 t0, double2 t1, double2 t2) {
    const double v1x = t1.x - t0.x, v1y = t1.y - t0.y;
    const double v2x = t2.x - t0.x, v2y = t2.y - t0.y;
    const double d12 = det2(v1x, v1y, v2x, v2y);
    const double a = (det2(px, py, v2x, v2y) - det2(t0.x, t0.y, v2x, v2y)) / d12;
    const double b = -(det2(px, py, v1x, v1y) - det2(t0.x, t0.y, v1x, v1y)) / d12;
    return a >= 0.0 && b >= 0.0 && a + b <= 1.0;
}

// points: grid mode (ny > 0): point idx = (i, j) -> (xa[i], ya[j]), idx = i*ny + j; list mode (ny == 0): (xa[idx], ya[idx])
__device__ __forceinline__ void point_of(int64_t idx, const double* xa, const double* ya, int64_t ny, double& x, double& y) {
    if (ny > 0) { x = xa[idx / ny]; y = ya[idx % ny]; }
    else { x = xa[idx]; y = ya[idx]; }
}

__global__ void k_locate(int64_t npts, const double* __restrict__ xa, const double* __restrict__ ya, int64_t ny, double maxr,
                         int tie_rule, Grid g, const int32_t* __restrict__ cell_ptr, const int32_t* __restrict__ cell_tri,
                         const int32_t* __restrict__ conn, int stride, const double2* __restrict__ xy, double bx1, double by1,
                         int64_t* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= npts) return;
    double x, y;
    point_of(idx, xa, ya, ny, x, y);
    int64_t best = -1;
    double bestd = 0.0;
    const bool far = maxr > 0.0 && sqrt(x * x + y * y) > maxr;                 // fe_solver.py:849-851
    if (!far && x >= g.x0 && y >= g.y0 && x <= bx1 && y <= by1) {
        const int cell = cell_y(g, y) * g.nx + cell_x(g, x);
        for (int s = cell_ptr[cell]; s < cell_ptr[cell + 1]; ++s) {
            const int32_t e = cell_tri[s];
            const int32_t* t = conn + (int64_t)e * stride;
            const double2 a = xy[t[0]], b = xy[t[1]], c = xy[t[2]];
            if (!isinside(x, y, a, b, c)) continue;
            if (tie_rule == 1) {                                               // first in mesh order (find_triangle)
                if (best < 0 || e < best) best = e;
            } else {                                                           // nearest centroid (KD-tree search order)
                const double cx = ((a.x + b.x) + c.x) / 3.0, cy = ((a.y + b.y) + c.y) / 3.0;   // mesher.py:239 np.mean
                const double d = (x - cx) * (x - cx) + (y - cy) * (y - cy);
                if (best < 0 || d < bestd || (d == bestd && e < best)) { best = e; bestd = d; }
            }
        }
    }
    out[idx] = best;
}

// kind 2: P2 (6 weights), kind 1: P1 (3 weights), kind 3: Whitney edge (3 x 2 weights)
__global__ void k_weights(int64_t npts, const double* __restrict__ xa, const double* __restrict__ ya, int64_t ny, int kind,
                          const int64_t* __restrict__ tri_idx, int64_t Ne, const int32_t* __restrict__ conn, int stride,
                          const double2* __restrict__ xy, double* __restrict__ w) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= npts) return;
    const int nw = kind == 2 ? 6 : (kind == 1 ? 3 : 6);
    double* o = w + idx * nw;
    const int64_t e = tri_idx[idx];
    if (e == -1) {
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        for (int k = 0; k < nw; ++k) o[k] = nan;                              // fe_solver.py:685-687, 710-712
        return;
    }
    const int64_t ee = e < 0 ? e + Ne : e;                                    // numpy negative indexing (other than -1)
    double x, y;
    point_of(idx, xa, ya, ny, x, y);
    const int32_t* t = conn + ee * stride;
    const int32_t i0 = t[0], i1 = t[1], i2 = t[2];
    const double2 P0 = xy[i0], P1 = xy[i1], P2 = xy[i2];
    const double x21 = P1.x - P0.x, y21 = P1.y - P0.y;
    const double x31 = P2.x - P0.x, y31 = P2.y - P0.y;
    const double J = x21 * y31 - x31 * y21;
    if (kind == 3) {
        // shape_funcs.py:331-353 (edge k: 12, 23, 31), signs shape_funcs.py:317-322
        const double x32 = P2.x - P1.x, y32 = P2.y - P1.y;
        const double l12 = sqrt(x21 * x21 + y21 * y21) * (i0 < i1 ? 1.0 : -1.0);
        const double l23 = sqrt(x32 * x32 + y32 * y32) * (i1 < i2 ? 1.0 : -1.0);
        const double l31 = sqrt(x31 * x31 + y31 * y31) * (i2 < i0 ? 1.0 : -1.0);
        o[0] = (((-y + y31) + P0.y) / J) * l12;
        o[1] = (((x - x31) - P0.x) / J) * l12;
        o[2] = ((-y + P0.y) / J) * l23;
        o[3] = ((x - P0.x) / J) * l23;
        o[4] = (((-y + y21) + P0.y) / J) * l31;
        o[5] = (((x - x21) - P0.x) / J) * l31;
        return;
    }
    // shape_funcs.py:36-49: M = [[y31, -x31], [-y21, x21]] / J ; uv = M (xy - P0)
    const double m00 = y31 / J, m01 = (-x31) / J, m10 = (-y21) / J, m11 = x21 / J;
    const double dx = x - P0.x, dy = y - P0.y;
    const double u = m00 * dx + m01 * dy, v = m10 * dx + m11 * dy;
    const double l = (1.0 - u) - v;
    if (kind == 2) {                                                          // shape_funcs.py:5-10
        o[0] = (2.0 * l) * ((0.5 - u) - v);
        o[1] = (2.0 * u) * (u - 0.5);
        o[2] = (2.0 * v) * (v - 0.5);
        o[3] = (4.0 * u) * l;
        o[4] = (4.0 * u) * v;
        o[5] = (4.0 * v) * l;
    } else {                                                                  // shape_funcs.py:286-288
        o[0] = l;
        o[1] = u;
        o[2] = v;
    }
}

// out[idx, c] = sum_k field[dof[e, k]] * w[idx, k, c]   (ncomp doubles per field entry: 1 real, 2 complex)
__global__ void k_apply(int64_t npts, const int64_t* __restrict__ tri_idx, int64_t Ne, const int32_t* __restrict__ dof, int stride,
                        int nd, int wdim, const double* __restrict__ w, const double* __restrict__ field, int ncomp,
                        double* __restrict__ out) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= npts) return;
    int64_t e = tri_idx[idx];
    if (e < 0) e += Ne;                          // field_points = v[tris[tri_idxs]]: -1 reads the last triangle (times NaN)
    const int32_t* d = dof + e * stride;
    const double* wi = w + idx * nd * wdim;
    for (int c = 0; c < wdim; ++c)
        for (int q = 0; q < ncomp; ++q) {
            double acc = 0.0;
            for (int k = 0; k < nd; ++k) {
                const double f = field[(int64_t)d[k] * ncomp + q];
                const double p = f * wi[k * wdim + c];
                acc = k == 0 ? p : acc + p;      // np.sum over 3 / 6 terms: sequential
            }
            out[(idx * wdim + c) * ncomp + q] = acc;
        }
}

}  // namespace

}  // namespace ws

using namespace ws;

extern "C" {

int ws_locator_create(const double* xy, const int32_t* conn, int64_t Ne, int stride, double bx0, double by0, double bx1,
                      double by1, ws_locator** out, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(out, "out == NULL");
    *out = nullptr;
    WS_REQUIRE(xy && conn && Ne > 0 && stride >= 3, "bad arguments");
    WS_REQUIRE(bx1 > bx0 && by1 > by0, "empty bounding box");
    DeviceGuard g(device);
    configure_mempool(device);
    cudaStream_t st = (cudaStream_t)stream;
    auto* L = new ws_locator();
    try {
        L->device = device;
        L->Ne = Ne;
        // about two bins per triangle, square bins
        const double W = bx1 - bx0, H = by1 - by0;
        const double cells = std::min(2.0 * (double)Ne, 32.0e6);
        L->nx = (int)std::max(1.0, std::ceil(std::sqrt(cells * W / H)));
        L->ny = (int)std::max(1.0, std::ceil(std::sqrt(cells * H / W)));
        L->x0 = bx0;
        L->y0 = by0;
        L->x1 = bx1;
        L->y1 = by1;
        L->inv_hx = L->nx / W;
        L->inv_hy = L->ny / H;
        const int64_t nc = (int64_t)L->nx * L->ny;
        Grid gr{L->nx, L->ny, L->x0, L->y0, L->inv_hx, L->inv_hy};
        DevBuf<int32_t> count(nc + 1, st);
        L->cell_ptr.alloc(nc + 1, st);
        WS_CUDA(cudaMemsetAsync(count.get(), 0, count.bytes(), st));
        const int T = 256;
        k_bin_tris<<<ceil_div(Ne, T), T, 0, st>>>(0, Ne, conn, stride, (const double2*)xy, gr, count.get(), nullptr, nullptr);
        size_t bytes = 0;
        WS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, count.get(), L->cell_ptr.get(), nc + 1, st));
        DevBuf<char> tmp(bytes, st);
        WS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, count.get(), L->cell_ptr.get(), nc + 1, st));
        int32_t total = 0;
        WS_CUDA(cudaMemcpyAsync(&total, L->cell_ptr.get() + nc, sizeof total, cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaStreamSynchronize(st));
        WS_REQUIRE(total >= 0, "bin lists exceed 32-bit indexing");
        L->cell_tri.alloc(std::max<int64_t>(total, 1), st);
        WS_CUDA(cudaMemsetAsync(count.get(), 0, count.bytes(), st));
        k_bin_tris<<<ceil_div(Ne, T), T, 0, st>>>(1, Ne, conn, stride, (const double2*)xy, gr, count.get(), L->cell_ptr.get(),
                                                  L->cell_tri.get());
        WS_CUDA(cudaStreamSynchronize(st));     // `count` / `tmp` die here
        WS_CUDA(cudaGetLastError());
        count_launch(2);
    } catch (...) {
        delete L;
        throw;
    }
    *out = L;
    WS_API_END
}

int ws_locator_destroy(ws_locator* L) {
    WS_API_BEGIN
    if (L) {
        DeviceGuard g(L->device);
        delete L;
    }
    WS_API_END
}

int ws_locate(const ws_locator* L, const double* xy, const int32_t* conn, int stride, const double* xa, const double* ya,
              int64_t nx, int64_t ny, double maxr, int tie_rule, int64_t* tri_idx, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(L && xy && conn && tri_idx, "null pointer");
    WS_REQUIRE(nx >= 0 && ny >= 0 && (tie_rule == 0 || tie_rule == 1), "bad arguments");
    const int64_t npts = ny > 0 ? nx * ny : nx;
    if (npts == 0) return WS_OK;
    WS_REQUIRE(xa && ya, "null pointer");
    DeviceGuard g(device);
    Grid gr{L->nx, L->ny, L->x0, L->y0, L->inv_hx, L->inv_hy};
    k_locate<<<ceil_div(npts, 128), 128, 0, (cudaStream_t)stream>>>(npts, xa, ya, ny, maxr, tie_rule, gr, L->cell_ptr.get(),
                                                                   L->cell_tri.get(), conn, stride, (const double2*)xy,
                                                                   L->x1, L->y1, tri_idx);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
    WS_API_END
}

int ws_interp_weights(const double* xy, const int32_t* conn, int64_t Ne, int stride, int kind, const double* xa,
                      const double* ya, int64_t nx, int64_t ny, const int64_t* tri_idx, double* weights, int device,
                      void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(kind == 1 || kind == 2 || kind == 3, "kind: 1 = P1, 2 = P2, 3 = Whitney edge");
    const int64_t npts = ny > 0 ? nx * ny : nx;
    if (npts == 0) return WS_OK;
    WS_REQUIRE(xy && conn && xa && ya && tri_idx && weights && Ne > 0, "null pointer");
    DeviceGuard g(device);
    k_weights<<<ceil_div(npts, 128), 128, 0, (cudaStream_t)stream>>>(npts, xa, ya, ny, kind, tri_idx, Ne, conn, stride,
                                                                    (const double2*)xy, weights);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
    WS_API_END
}

int ws_interp_apply(const int64_t* tri_idx, int64_t npts, int64_t Ne, const int32_t* dof, int stride, int nd, int wdim,
                    const double* weights, const double* field, int ncomp, double* out, int device, void* stream) {
    WS_API_BEGIN
    if (npts == 0) return WS_OK;
    WS_REQUIRE(tri_idx && dof && weights && field && out && Ne > 0, "null pointer");
    WS_REQUIRE((nd == 3 || nd == 6) && (wdim == 1 || wdim == 2) && (ncomp == 1 || ncomp == 2) && stride >= nd, "bad shape");
    DeviceGuard g(device);
    k_apply<<<ceil_div(npts, 128), 128, 0, (cudaStream_t)stream>>>(npts, tri_idx, Ne, dof, stride, nd, wdim, weights, field,
                                                                  ncomp, out);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
    WS_API_END
}

}  // extern "C"
