// K3 / K4: FEM element integration + assembly into the CSR pattern, fp64.
//
// Design (B200-first, not the reference's COO-triplet flow): the matrix is assembled by
// ROWS.  The pattern object holds, for every matrix row, the list of (element, local index)
// incidences sorted by element -- the reference's summation order -- and the position of
// each of the element's six columns inside that row.  One thread owns one row, computes for
// each incident element only the row of the 6x6 element matrices it needs, and accumulates
// into a shared-memory tile that covers the CTA's contiguous slot range; the tile is then
// written with coalesced 8-byte stores.  Every output value is written exactly once, there
// are no atomics and no zero-fill pass, and the result is deterministic.
//
// Arithmetic follows the reference expression by expression (same association, true
// divisions, separate multiply and subtract: this file is compiled with -fmad=false), so
// element entries are bit-identical to numpy's and only the order in which scipy sums
// duplicates can differ (scalar path); the vector path sums in the reference's element
// order and reproduces the entries that cancel to exactly 0.0.
//
//   scalar P2 : shape_funcs.py:192-246 (compute_NN_dNdN_vec), fe_solver.py:51-54
//   vector    : shape_funcs.py:310-428 (precompute, computeL_*), fe_solver.py:196-199
#include "ws_common.cuh"

namespace ws {

constexpr int ROWS_PER_CTA = 128;
constexpr int TILE_CAP = 3072;  // slots per CTA tile in shared memory (A and B: 48 KB -> 4 CTAs / SM);
                                // denser row blocks fall back to accumulating in global memory

struct RowCtx {
    const int32_t* rowptr;
    const int32_t* incptr;
    const uint32_t* inc;
    const uint16_t* pos;
    int64_t N;
};

// ---- scalar P2 ---------------------------------------------------------------------------
// Row `a` of the element matrices, using the cyclic symmetry of the P2 triangle: rotate the
// vertices so that the row becomes row 0 (vertex) or row 3 (mid-side).  Rotation only negates
// or renames the coordinate differences, and round-to-nearest is sign-symmetric, so each
// entry has the same bits as the reference's un-rotated formula.  J is taken from the
// un-rotated vertices (shape_funcs.py:202).
struct ScalarP2 {
    const double2* xy;
    const int32_t* dofs;
    const double* k2n2;   // nullptr -> mass only

    __device__ __forceinline__ void row(uint32_t e, int a, double (&Av)[6], double (&Bv)[6], int (&cb)[6]) const {
        const int32_t* d = dofs + (int64_t)e * 6;
        double2 P0 = xy[d[0]], P1 = xy[d[1]], P2 = xy[d[2]];
        const double J = (P1.x - P0.x) * (P2.y - P0.y) - (P2.x - P0.x) * (P1.y - P0.y);
        const int rot = a < 3 ? a : a - 3;
        if (rot == 1) { double2 t = P0; P0 = P1; P1 = P2; P2 = t; }
        else if (rot == 2) { double2 t = P0; P0 = P2; P2 = P1; P1 = t; }
        const double x21 = P1.x - P0.x, y21 = P1.y - P0.y;
        const double x31 = P2.x - P0.x, y31 = P2.y - P0.y;
        const double x32 = P2.x - P1.x, y32 = P2.y - P1.y;
        // new-local column j -> original local column
        const int r1 = rot == 2 ? 0 : rot + 1, r2 = rot == 0 ? 2 : rot - 1;
        cb[0] = rot; cb[1] = r1; cb[2] = r2; cb[3] = 3 + rot; cb[4] = 3 + r1; cb[5] = 3 + r2;
        double K[6], M[6];
        if (a < 3) {
            const double s01 = y32 * y31 + x32 * x31;
            const double s02 = y21 * y32 + x21 * x32;
            K[0] = (y32 * y32 + x32 * x32) / (2 * J);
            K[1] = s01 / (6 * J);
            K[2] = (-y21 * y32 - x21 * x32) / (6 * J);
            K[3] = -2 * s01 / (3 * J);
            K[4] = 0.0;
            K[5] = 2 * s02 / (3 * J);
            M[0] = J / 60;
            M[1] = M[2] = -J / 360;
            M[3] = M[5] = 0.0;
            M[4] = -J / 90;
        } else {
            const double f = 4 / (3 * J);
            const double s01 = y32 * y31 + x32 * x31;
            K[0] = K[1] = -2 * s01 / (3 * J);
            K[2] = 0.0;
            K[3] = f * (y32 * y32 + y31 * y21 + x32 * x32 + x31 * x21);
            K[4] = f * (y21 * y32 + x21 * x32);
            K[5] = -f * (y31 * y21 + x31 * x21);
            M[0] = M[1] = 0.0;
            M[2] = -J / 90;
            M[3] = 4 * J / 45;
            M[4] = M[5] = 2 * J / 45;
        }
        if (k2n2) {
            const double c = k2n2[e];
#pragma unroll
            for (int j = 0; j < 6; ++j) Av[j] = __dsub_rn(__dmul_rn(c, M[j]), K[j]);   // fe_solver.py:51
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) Bv[j] = M[j];                                        // fe_solver.py:54
    }
};

// ---- vector: Whitney edge elements + P1 nodes ---------------------------------------------
struct VectorP1 {
    const double2* xy;
    const int32_t* tris;
    const double* k2n2;
    int transpose;
    int qd;          // 1: write the quasi-definite shift-invert matrix M' = T^T (A - sigma B) T into Bv
    double sigma;

    // Branch-free in the local index: the (cheap) numerators of all rows are formed and the ones
    // of row `a` are picked with selects, so a warp whose lanes hold different local rows does not
    // serialise three copies of the expensive part (3 sqrt + 7-8 IEEE divisions per incidence).
    // Every expression keeps the reference's association (shape_funcs.py line cited per entry).
    __device__ __forceinline__ void row(uint32_t e, int a, double (&Av)[6], double (&Bv)[6], int (&cb)[6]) const {
        const int32_t* t = tris + (int64_t)e * 3;
        const int32_t t0 = t[0], t1 = t[1], t2 = t[2];
        const double2 P0 = xy[t0], P1 = xy[t1], P2 = xy[t2];
        const double x21 = P1.x - P0.x, y21 = P1.y - P0.y;          // shape_funcs.py:311-316
        const double x31 = P2.x - P0.x, y31 = P2.y - P0.y;
        const double x32 = P2.x - P1.x, y32 = P2.y - P1.y;
        const double s1 = t0 < t1 ? 1.0 : -1.0;                      // shape_funcs.py:317-319
        const double s2 = t1 < t2 ? 1.0 : -1.0;
        const double s3 = t2 < t0 ? 1.0 : -1.0;
        const double J = x21 * y31 - x31 * y21;                      // shape_funcs.py:323
        const double c = k2n2[e];
        const double xx = x21 * x31, yy = y21 * y31;
        const double l12 = sqrt(x21 * x21 + y21 * y21) * s1;         // shape_funcs.py:320-322
        const double l23 = sqrt(x32 * x32 + y32 * y32) * s2;
        const double l31 = sqrt(x31 * x31 + y31 * y31) * s3;
        const double sq12 = l12 * l12, sq23 = l23 * l23, sq31 = l31 * l31;
        const double x3 = (3 * x21) * x31, y3 = (3 * y21) * y31;
        const double J6 = 6 * J;
#pragma unroll
        for (int j = 0; j < 6; ++j) { cb[j] = j; Av[j] = 0.0; }
        // numerators of NedN (D_ij, i = edge, j = node), shape_funcs.py:393-401
        const double nD00 = l12 * (-sq12 + x3 - 2 * sq31 + y3);       // :393
        const double nD01 = l12 * (-xx + 2 * sq31 - yy);              // :396
        const double nD02 = l12 * (sq12 - 2 * xx - 2 * yy);           // :401
        const double nD10 = l23 * (-sq12 + sq31);                     // :397
        const double nD11 = l23 * (-xx - sq31 - yy);                  // :394
        const double nD12 = l23 * (sq12 + xx + yy);                   // :398
        const double nD20 = l31 * (2 * sq12 - x3 + sq31 - y3);        // :400
        const double nD21 = l31 * (2 * xx + 2 * yy - sq31);           // :399
        const double nD22 = l31 * (-2 * sq12 + xx + yy);              // :395
        if (a < 3) {
            const double J12 = 12 * J, tJ = 2 / J;
            // numerators of NeNe (symmetric), shape_funcs.py:369-374
            const double nE00 = sq12 * (sq12 - x3 + 3 * sq31 - y3);          // :369
            const double nE11 = sq23 * (sq12 + xx + sq31 + yy);              // :370
            const double nE22 = sq31 * (3 * sq12 - x3 + sq31 - y3);          // :371
            const double nE01 = l12 * l23 * (sq12 - xx - sq31 - yy);         // :372
            const double nE12 = l23 * l31 * (-sq12 - xx + sq31 - yy);        // :373
            const double nE02 = l31 * l12 * (-sq12 + x3 - sq31 + y3);        // :374
            const bool a0 = a == 0, a1 = a == 1;
            const double E0 = (a0 ? nE00 : (a1 ? nE01 : nE02)) / J12;
            const double E1 = (a0 ? nE01 : (a1 ? nE11 : nE12)) / J12;
            const double E2 = (a0 ? nE02 : (a1 ? nE12 : nE22)) / J12;
            const double D0 = (a0 ? nD00 : (a1 ? nD10 : nD20)) / J6;
            const double D1 = (a0 ? nD01 : (a1 ? nD11 : nD21)) / J6;
            const double D2 = (a0 ? nD02 : (a1 ? nD12 : nD22)) / J6;
            const double la = a0 ? l12 : (a1 ? l23 : l31);
            // curl-curl ((2/J) l_i) l_j, shape_funcs.py:382-383 (row i = a; transposed: column i = a)
            const double C0 = transpose ? (tJ * l12) * la : (tJ * la) * l12;
            const double C1 = transpose ? (tJ * l23) * la : (tJ * la) * l23;
            const double C2 = transpose ? (tJ * l31) * la : (tJ * la) * l31;
            if (qd) {
                Bv[0] = (c - sigma) * E0 - C0; Bv[1] = (c - sigma) * E1 - C1; Bv[2] = (c - sigma) * E2 - C2;
                Bv[3] = -c * D0; Bv[4] = -c * D1; Bv[5] = -c * D2;
                return;
            }
            Av[0] = __dsub_rn(__dmul_rn(c, E0), C0);                              // fe_solver.py:196
            Av[1] = __dsub_rn(__dmul_rn(c, E1), C1);
            Av[2] = __dsub_rn(__dmul_rn(c, E2), C2);
            Bv[0] = E0; Bv[1] = E1; Bv[2] = E2;                                    // fe_solver.py:197
            Bv[3] = D0; Bv[4] = D1; Bv[5] = D2;                                    // fe_solver.py:198
        } else {
            const int i = a - 3;
            const bool i0 = i == 0, i1 = i == 1;
            const double J2 = 2 * J;
            const double Nd = J / 12, No = J / 24;                                // shape_funcs.py:410-411
            // numerators of dNdN (symmetric), shape_funcs.py:421-426
            const double nG00 = sq12 + sq31 - 2 * xx - 2 * yy;                    // :421
            const double nG11 = sq31;                                             // :422
            const double nG22 = sq12;                                             // :423
            const double nG01 = -sq31 + xx + yy;                                  // :424
            const double nG12 = -xx - yy;                                         // :425
            const double nG02 = -sq12 + xx + yy;                                  // :426
            const double Dt0 = (i0 ? nD00 : (i1 ? nD01 : nD02)) / J6;             // column i of NedN
            const double Dt1 = (i0 ? nD10 : (i1 ? nD11 : nD12)) / J6;
            const double Dt2 = (i0 ? nD20 : (i1 ? nD21 : nD22)) / J6;
            const double G0 = (i0 ? nG00 : (i1 ? nG01 : nG02)) / J2;
            const double G1 = (i0 ? nG01 : (i1 ? nG11 : nG12)) / J2;
            const double G2 = (i0 ? nG02 : (i1 ? nG12 : nG22)) / J2;
            const double N0 = i0 ? Nd : No, N1 = i1 ? Nd : No, N2 = (i == 2) ? Nd : No;
            if (qd) {
                Bv[0] = -c * Dt0; Bv[1] = -c * Dt1; Bv[2] = -c * Dt2;
                Bv[3] = c * (G0 + sigma * N0); Bv[4] = c * (G1 + sigma * N1); Bv[5] = c * (G2 + sigma * N2);
                return;
            }
            Bv[0] = Dt0; Bv[1] = Dt1; Bv[2] = Dt2;                                // fe_solver.py:209
            Bv[3] = __dsub_rn(G0, __dmul_rn(c, N0));                              // fe_solver.py:199
            Bv[4] = __dsub_rn(G1, __dmul_rn(c, N1));
            Bv[5] = __dsub_rn(G2, __dmul_rn(c, N2));
        }
    }
};

template <class Op, bool WRITE_A>
__global__ void __launch_bounds__(ROWS_PER_CTA)
k_assemble_rows(RowCtx cx, Op op, double* __restrict__ A, double* __restrict__ B) {
    extern __shared__ double tile[];
    const int64_t r0 = (int64_t)blockIdx.x * ROWS_PER_CTA;
    const int64_t r1 = min(r0 + ROWS_PER_CTA, cx.N);
    const int s0 = cx.rowptr[r0], s1 = cx.rowptr[r1];
    const int S = s1 - s0;
    const bool in_smem = S <= TILE_CAP;
    double* tB = in_smem ? tile : B + s0;
    double* tA = in_smem ? tile + TILE_CAP : (WRITE_A ? A + s0 : nullptr);
    for (int t = threadIdx.x; t < S; t += ROWS_PER_CTA) {
        tB[t] = 0.0;
        if (WRITE_A) tA[t] = 0.0;
    }
    __syncthreads();   // (global fallback: the zero-fill above is visible CTA-wide after the barrier)
    const int64_t r = r0 + threadIdx.x;
    if (r < r1) {
        const int base = cx.rowptr[r] - s0;
        const int i0 = cx.incptr[r], i1 = cx.incptr[r + 1];
        for (int i = i0; i < i1; ++i) {
            const uint32_t u = cx.inc[i];
            double Av[6], Bv[6];
            int cb[6];
            op.row(u >> 3, (int)(u & 7u), Av, Bv, cb);
            const uint16_t* p = cx.pos + (int64_t)i * 6;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int q = base + p[cb[j]];
                tB[q] += Bv[j];
                if (WRITE_A) tA[q] += Av[j];
            }
        }
    }
    if (in_smem) {
        __syncthreads();
        for (int t = threadIdx.x; t < S; t += ROWS_PER_CTA) {
            B[s0 + t] = tB[t];
            if (WRITE_A) A[s0 + t] = tA[t];
        }
    }
}

template <class Op, bool WRITE_A>
static void launch_rows(const ws_pattern* p, const Op& op, double* A, double* B, cudaStream_t st) {
    if (p->N == 0 || p->nnz == 0) return;
    RowCtx cx{p->rowptr.get(), p->incptr.get(), p->inc.get(), p->pos.get(), p->N};
    size_t smem = (size_t)TILE_CAP * 8 * (WRITE_A ? 2 : 1);
    auto kern = k_assemble_rows<Op, WRITE_A>;
    static bool configured = false;
    if (!configured) {
        WS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    kern<<<ceil_div(p->N, ROWS_PER_CTA), ROWS_PER_CTA, smem, st>>>(cx, op, A, B);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
}

__global__ void k_axpby(int64_t n, double alpha, const double* __restrict__ x, double beta, double* __restrict__ y) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) y[i] = __dadd_rn(__dmul_rn(alpha, x[i]), __dmul_rn(beta, y[i]));
}

}  // namespace ws

using namespace ws;

extern "C" {

int ws_assemble_scalar_p2(const ws_pattern* p, const double* xy, const int32_t* dofs,
                          const double* elem_k2n2, double* A_vals, double* B_vals, int transpose,
                          int device, void* stream) {
    WS_API_BEGIN
    (void)transpose;   // P2 scalar element matrices are bit-symmetric
    WS_REQUIRE(p, "pattern == NULL");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && dofs && elem_k2n2 && A_vals && B_vals, "null pointer");
    DeviceGuard g(device);
    ScalarP2 op{(const double2*)xy, dofs, elem_k2n2};
    launch_rows<ScalarP2, true>(p, op, A_vals, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_mass_p2(const ws_pattern* p, const double* xy, const int32_t* dofs, double* B_vals,
                        int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && dofs && B_vals, "null pointer");
    DeviceGuard g(device);
    ScalarP2 op{(const double2*)xy, dofs, nullptr};
    launch_rows<ScalarP2, false>(p, op, nullptr, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_vector_p1(const ws_pattern* p, const double* xy, const int32_t* tris,
                          const double* elem_k2n2, int64_t Ntt, double* A_vals, double* B_vals,
                          int transpose, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    WS_REQUIRE(Ntt >= 0 && Ntt <= p->N, "0 <= Ntt <= N");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && tris && elem_k2n2 && A_vals && B_vals, "null pointer");
    DeviceGuard g(device);
    VectorP1 op{(const double2*)xy, tris, elem_k2n2, transpose, 0, 0.0};
    launch_rows<VectorP1, true>(p, op, A_vals, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_vector_qd(const ws_pattern* p, const double* xy, const int32_t* tris,
                          const double* elem_k2n2, int64_t Ntt, double sigma, double* M_vals,
                          int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    WS_REQUIRE(Ntt >= 0 && Ntt <= p->N, "0 <= Ntt <= N");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && tris && elem_k2n2 && M_vals, "null pointer");
    DeviceGuard g(device);
    VectorP1 op{(const double2*)xy, tris, elem_k2n2, 0, 1, sigma};
    launch_rows<VectorP1, false>(p, op, nullptr, M_vals, (cudaStream_t)stream);
    WS_API_END
}

// x = T y (mode 0):  x_t[i] = y_t[i] - (y_z[hi] - y_z[lo]) / len_i ,  x_z = y_z
// y = T^T b (mode 1): y_t = b_t ,  y_z[j] = b_z[j] - sum_{i: hi=j} b_t[i]/len_i + sum_{i: lo=j} b_t[i]/len_i
// The node rows gather over the pattern row (which lists every edge of the surrounding
// triangles) so the sum order is fixed.
__global__ void k_vector_T(int64_t N, int64_t Ntt, int mode, const int32_t* __restrict__ rowptr,
                           const int32_t* __restrict__ colidx, const int2* __restrict__ edges,
                           const double2* __restrict__ xy, const double* __restrict__ in, double* __restrict__ out) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    if (r < Ntt) {
        double v = in[r];
        if (mode == 0) {
            const int2 e = edges[r];
            const double2 a = xy[e.x], b = xy[e.y];
            const double len = sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
            v -= (in[Ntt + e.y] - in[Ntt + e.x]) / len;
        }
        out[r] = v;
    } else {
        double v = in[r];
        if (mode == 1) {
            const int j = (int)(r - Ntt);
            for (int s = rowptr[r]; s < rowptr[r + 1]; ++s) {
                const int c = colidx[s];
                if (c >= Ntt) break;            // columns are sorted: edges first
                const int2 e = edges[c];
                if (e.x != j && e.y != j) continue;
                const double2 a = xy[e.x], b = xy[e.y];
                const double len = sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
                v += (e.y == j ? -in[c] : in[c]) / len;
            }
        }
        out[r] = v;
    }
}

int ws_vector_T(const ws_pattern* p, const int32_t* edges, const double* xy, int64_t Ntt, int mode,
                const double* in, double* out, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p && edges && xy && in && out, "null pointer");
    WS_REQUIRE(in != out, "in-place application is not supported");
    WS_REQUIRE(Ntt >= 0 && Ntt <= p->N && (mode == 0 || mode == 1), "bad Ntt / mode");
    DeviceGuard g(device);
    if (p->N) {
        k_vector_T<<<ceil_div(p->N, 256), 256, 0, (cudaStream_t)stream>>>(p->N, Ntt, mode, p->rowptr.get(), p->colidx.get(),
                                                                          (const int2*)edges, (const double2*)xy, in, out);
        WS_CUDA(cudaGetLastError());
        count_launch(1);
    }
    WS_API_END
}

int ws_axpby(int64_t n, double alpha, const double* x, double beta, double* y, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(n >= 0, "n >= 0");
    if (n == 0) return WS_OK;
    WS_REQUIRE(x && y, "null pointer");
    DeviceGuard g(device);
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 16);
    k_axpby<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, alpha, x, beta, y);
    WS_CUDA(cudaGetLastError());
    count_launch(1);
    WS_API_END
}

}  // extern "C"
