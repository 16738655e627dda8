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
#include <algorithm>
#include <cstdlib>

#include <cub/cub.cuh>

#include "ws_common.cuh"

namespace ws {

constexpr int ROWS_PER_CTA = 128;
constexpr int MAX_TILE_BYTES = 96 * 1024;   // larger row blocks accumulate in global memory instead

struct RowCtx {
    const int32_t* rowptr;
    const int32_t* incptr;
    const uint32_t* inc;
    const uint16_t* pos;
    int64_t N;
    int64_t first_block;  // first 128-row block of this launch
    int cap;              // slots of the shared-memory tile (>= the largest block of the launch, if it fits)
    int batch;            // 1 (default): the six slots of an incidence are read in one pass and written in a second (WS_ASM_BATCH=0: one by one)
};

// raw inputs of one element: vertex ids, vertex coordinates, k^2 n^2
struct RawTri {
    int32_t t0, t1, t2;
    double2 P0, P1, P2;
    double c;
};

// ---- scalar P2 ---------------------------------------------------------------------------
// Row `a` of the element matrices, using the cyclic symmetry of the P2 triangle: rotate the
// vertices so that the row becomes row 0 (vertex) or row 3 (mid-side).  Rotation only negates
// or renames the coordinate differences, and round-to-nearest is sign-symmetric, so each
// entry has the same bits as the reference's un-rotated formula.  J is taken from the
// un-rotated vertices (shape_funcs.py:202).
// fast != 0: one reciprocal instead of the 5-8 IEEE divisions per row (entries within 2 ulp of
// the exact mode); used inside the eigen-solve pipeline, never for the matrices handed back to
// the caller.
struct ScalarP2 {
    const double2* xy;
    const int32_t* dofs;
    const double* k2n2;   // nullptr -> mass only
    int fast;

    __device__ __forceinline__ RawTri load(uint32_t e) const {
        const int32_t* d = dofs + (int64_t)e * 6;
        RawTri r;
        r.t0 = d[0]; r.t1 = d[1]; r.t2 = d[2];
        r.P0 = xy[r.t0]; r.P1 = xy[r.t1]; r.P2 = xy[r.t2];
        r.c = k2n2 ? k2n2[e] : 0.0;
        return r;
    }

    __device__ __forceinline__ void row(const RawTri& in, int a, double (&Av)[6], double (&Bv)[6], int (&cb)[6]) const {
        double2 P0 = in.P0, P1 = in.P1, P2 = in.P2;
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
        const double s01 = y32 * y31 + x32 * x31;
        if (!fast) {
            if (a < 3) {
                const double s02 = y21 * y32 + x21 * x32;
                K[0] = (y32 * y32 + x32 * x32) / (2 * J);          // shape_funcs.py:222-224
                K[1] = s01 / (6 * J);                              // :230-232
                K[2] = (-y21 * y32 - x21 * x32) / (6 * J);
                K[3] = -2 * s01 / (3 * J);                         // :235-237
                K[4] = 0.0;
                K[5] = 2 * s02 / (3 * J);
                M[0] = J / 60;                                     // :205
                M[1] = M[2] = -J / 360;                            // :209
                M[3] = M[5] = 0.0;
                M[4] = -J / 90;                                    // :212
            } else {
                const double f = 4 / (3 * J);
                K[0] = K[1] = -2 * s01 / (3 * J);                  // :235-237
                K[2] = 0.0;
                K[3] = f * (y32 * y32 + y31 * y21 + x32 * x32 + x31 * x21);   // :225-227
                K[4] = f * (y21 * y32 + x21 * x32);                // :242-244
                K[5] = -f * (y31 * y21 + x31 * x21);
                M[0] = M[1] = 0.0;
                M[2] = -J / 90;                                    // :212
                M[3] = 4 * J / 45;                                 // :206
                M[4] = M[5] = 2 * J / 45;                          // :214
            }
        } else {
            const double rJ = 1.0 / J;
            if (a < 3) {
                const double s02 = y21 * y32 + x21 * x32;
                K[0] = (y32 * y32 + x32 * x32) * (0.5 * rJ);
                K[1] = s01 * (rJ * (1.0 / 6.0));
                K[2] = -s02 * (rJ * (1.0 / 6.0));
                K[3] = -2 * s01 * (rJ * (1.0 / 3.0));
                K[4] = 0.0;
                K[5] = 2 * s02 * (rJ * (1.0 / 3.0));
                M[0] = J * (1.0 / 60.0);
                M[1] = M[2] = -J * (1.0 / 360.0);
                M[3] = M[5] = 0.0;
                M[4] = -J * (1.0 / 90.0);
            } else {
                const double f = 4 * (rJ * (1.0 / 3.0));
                K[0] = K[1] = -2 * s01 * (rJ * (1.0 / 3.0));
                K[2] = 0.0;
                K[3] = f * (y32 * y32 + y31 * y21 + x32 * x32 + x31 * x21);
                K[4] = f * (y21 * y32 + x21 * x32);
                K[5] = -f * (y31 * y21 + x31 * x21);
                M[0] = M[1] = 0.0;
                M[2] = -J * (1.0 / 90.0);
                M[3] = 4 * J * (1.0 / 45.0);
                M[4] = M[5] = 2 * J * (1.0 / 45.0);
            }
        }
        if (k2n2) {
#pragma unroll
            for (int j = 0; j < 6; ++j) Av[j] = __dsub_rn(__dmul_rn(in.c, M[j]), K[j]);   // fe_solver.py:51
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) Bv[j] = M[j];                                        // fe_solver.py:54
    }
};

// ---- scalar P1 (fe_solver.py:103-129, shape_funcs.py:403-428) ------------------------------
// The pattern machinery works on six unknowns per element; a linear triangle is presented as
// (v0 v1 v2 v0 v1 v2): local rows 3..5 are aliases that contribute exact zeros, local rows 0..2
// carry LNN / LdNdN.  (order=1 scalar is not in any benchmark configuration.)
struct ScalarP1 {
    const double2* xy;
    const int32_t* dofs;
    const double* k2n2;

    __device__ __forceinline__ RawTri load(uint32_t e) const {
        const int32_t* d = dofs + (int64_t)e * 6;
        RawTri r;
        r.t0 = d[0]; r.t1 = d[1]; r.t2 = d[2];
        r.P0 = xy[r.t0]; r.P1 = xy[r.t1]; r.P2 = xy[r.t2];
        r.c = k2n2[e];
        return r;
    }

    __device__ __forceinline__ void row(const RawTri& in, int a, double (&Av)[6], double (&Bv)[6], int (&cb)[6]) const {
#pragma unroll
        for (int j = 0; j < 6; ++j) { cb[j] = j; Av[j] = 0.0; Bv[j] = 0.0; }
        if (a >= 3) return;
        const double2 P0 = in.P0, P1 = in.P1, P2 = in.P2;
        const double x21 = P1.x - P0.x, y21 = P1.y - P0.y;          // shape_funcs.py:311-314
        const double x31 = P2.x - P0.x, y31 = P2.y - P0.y;
        const double J = x21 * y31 - x31 * y21;                      // :323
        const double l12 = sqrt(x21 * x21 + y21 * y21), l31 = sqrt(x31 * x31 + y31 * y31);   // :320-322 (sign squares away)
        const double sq12 = l12 * l12, sq31 = l31 * l31;
        const double xx = x21 * x31, yy = y21 * y31;
        const double nG00 = sq12 + sq31 - 2 * xx - 2 * yy;          // :421
        const double nG11 = sq31, nG22 = sq12;                      // :422-423
        const double nG01 = -sq31 + xx + yy;                        // :424
        const double nG12 = -xx - yy;                               // :425
        const double nG02 = -sq12 + xx + yy;                        // :426
        const bool i0 = a == 0, i1 = a == 1;
        const double g0 = i0 ? nG00 : (i1 ? nG01 : nG02);
        const double g1 = i0 ? nG01 : (i1 ? nG11 : nG12);
        const double g2 = i0 ? nG02 : (i1 ? nG12 : nG22);
        const double J2 = 2 * J;
        const double G0 = g0 / J2, G1 = g1 / J2, G2 = g2 / J2;
        const double Nd = J / 12, No = J / 24;                      // :410-411
        const double N0 = i0 ? Nd : No, N1 = i1 ? Nd : No, N2 = (a == 2) ? Nd : No;
        Av[0] = __dsub_rn(__dmul_rn(in.c, N0), G0);                 // fe_solver.py:127
        Av[1] = __dsub_rn(__dmul_rn(in.c, N1), G1);
        Av[2] = __dsub_rn(__dmul_rn(in.c, N2), G2);
        Bv[0] = N0; Bv[1] = N1; Bv[2] = N2;                         // fe_solver.py:128
    }
};

// ---- vector: Whitney edge elements + P1 nodes ---------------------------------------------
// mode 0: Av = row of A, Bv = row of B (reference semantics, fe_solver.py:196-199)
// mode 1: Bv = row of the quasi-definite M' = T^T (A - sigma B) T
// mode 2: Av = row of M', Bv = row of B (what the eigen-solve pipeline needs, one pass)
struct VectorP1 {
    const double2* xy;
    const int32_t* tris;
    const double* k2n2;
    int transpose;
    int mode;
    double sigma;
    int fast;

    __device__ __forceinline__ RawTri load(uint32_t e) const {
        const int32_t* t = tris + (int64_t)e * 3;
        RawTri r;
        r.t0 = t[0]; r.t1 = t[1]; r.t2 = t[2];
        r.P0 = xy[r.t0]; r.P1 = xy[r.t1]; r.P2 = xy[r.t2];
        r.c = k2n2[e];
        return r;
    }

    // Branch-free in the local index: the (cheap) numerators of all rows are formed and the ones
    // of row `a` are picked with selects, so a warp whose lanes hold different local rows does not
    // serialise three copies of the expensive part (3 sqrt + 7-8 IEEE divisions per incidence).
    // Every expression keeps the reference's association (shape_funcs.py line cited per entry).
    __device__ __forceinline__ void row(const RawTri& in, int a, double (&Av)[6], double (&Bv)[6], int (&cb)[6]) const {
        const int32_t t0 = in.t0, t1 = in.t1, t2 = in.t2;
        const double2 P0 = in.P0, P1 = in.P1, P2 = in.P2;
        const double x21 = P1.x - P0.x, y21 = P1.y - P0.y;          // shape_funcs.py:311-316
        const double x31 = P2.x - P0.x, y31 = P2.y - P0.y;
        const double x32 = P2.x - P1.x, y32 = P2.y - P1.y;
        const double s1 = t0 < t1 ? 1.0 : -1.0;                      // shape_funcs.py:317-319
        const double s2 = t1 < t2 ? 1.0 : -1.0;
        const double s3 = t2 < t0 ? 1.0 : -1.0;
        const double J = x21 * y31 - x31 * y21;                      // shape_funcs.py:323
        const double c = in.c;
        const double xx = x21 * x31, yy = y21 * y31;
        const double l12 = sqrt(x21 * x21 + y21 * y21) * s1;         // shape_funcs.py:320-322
        const double l23 = sqrt(x32 * x32 + y32 * y32) * s2;
        const double l31 = sqrt(x31 * x31 + y31 * y31) * s3;
        const double sq12 = l12 * l12, sq23 = l23 * l23, sq31 = l31 * l31;
        const double x3 = (3 * x21) * x31, y3 = (3 * y21) * y31;
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
        const double rJ = fast ? 1.0 / J : 0.0;
        if (a < 3) {
            // numerators of NeNe (symmetric), shape_funcs.py:369-374
            const double nE00 = sq12 * (sq12 - x3 + 3 * sq31 - y3);          // :369
            const double nE11 = sq23 * (sq12 + xx + sq31 + yy);              // :370
            const double nE22 = sq31 * (3 * sq12 - x3 + sq31 - y3);          // :371
            const double nE01 = l12 * l23 * (sq12 - xx - sq31 - yy);         // :372
            const double nE12 = l23 * l31 * (-sq12 - xx + sq31 - yy);        // :373
            const double nE02 = l31 * l12 * (-sq12 + x3 - sq31 + y3);        // :374
            const bool a0 = a == 0, a1 = a == 1;
            const double e0 = a0 ? nE00 : (a1 ? nE01 : nE02);
            const double e1 = a0 ? nE01 : (a1 ? nE11 : nE12);
            const double e2 = a0 ? nE02 : (a1 ? nE12 : nE22);
            const double d0 = a0 ? nD00 : (a1 ? nD10 : nD20);
            const double d1 = a0 ? nD01 : (a1 ? nD11 : nD21);
            const double d2 = a0 ? nD02 : (a1 ? nD12 : nD22);
            double E0, E1, E2, D0, D1, D2, tJ;
            if (!fast) {
                const double J12 = 12 * J, J6 = 6 * J;
                E0 = e0 / J12; E1 = e1 / J12; E2 = e2 / J12;
                D0 = d0 / J6; D1 = d1 / J6; D2 = d2 / J6;
                tJ = 2 / J;
            } else {
                const double i12 = rJ * (1.0 / 12.0), i6 = rJ * (1.0 / 6.0);
                E0 = e0 * i12; E1 = e1 * i12; E2 = e2 * i12;
                D0 = d0 * i6; D1 = d1 * i6; D2 = d2 * i6;
                tJ = 2 * rJ;
            }
            const double la = a0 ? l12 : (a1 ? l23 : l31);
            // curl-curl ((2/J) l_i) l_j, shape_funcs.py:382-383 (row i = a; transposed: column i = a)
            const double C0 = transpose ? (tJ * l12) * la : (tJ * la) * l12;
            const double C1 = transpose ? (tJ * l23) * la : (tJ * la) * l23;
            const double C2 = transpose ? (tJ * l31) * la : (tJ * la) * l31;
            if (mode != 0) {
                double (&Mv)[6] = mode == 1 ? Bv : Av;
                Mv[0] = (c - sigma) * E0 - C0; Mv[1] = (c - sigma) * E1 - C1; Mv[2] = (c - sigma) * E2 - C2;
                Mv[3] = -c * D0; Mv[4] = -c * D1; Mv[5] = -c * D2;
                if (mode == 1) return;
            } else {
                Av[0] = __dsub_rn(__dmul_rn(c, E0), C0);                          // fe_solver.py:196
                Av[1] = __dsub_rn(__dmul_rn(c, E1), C1);
                Av[2] = __dsub_rn(__dmul_rn(c, E2), C2);
            }
            Bv[0] = E0; Bv[1] = E1; Bv[2] = E2;                                    // fe_solver.py:197
            Bv[3] = D0; Bv[4] = D1; Bv[5] = D2;                                    // fe_solver.py:198
        } else {
            const int i = a - 3;
            const bool i0 = i == 0, i1 = i == 1;
            // numerators of dNdN (symmetric), shape_funcs.py:421-426
            const double nG00 = sq12 + sq31 - 2 * xx - 2 * yy;                    // :421
            const double nG11 = sq31;                                             // :422
            const double nG22 = sq12;                                             // :423
            const double nG01 = -sq31 + xx + yy;                                  // :424
            const double nG12 = -xx - yy;                                         // :425
            const double nG02 = -sq12 + xx + yy;                                  // :426
            const double d0 = i0 ? nD00 : (i1 ? nD01 : nD02);                     // column i of NedN
            const double d1 = i0 ? nD10 : (i1 ? nD11 : nD12);
            const double d2 = i0 ? nD20 : (i1 ? nD21 : nD22);
            const double g0 = i0 ? nG00 : (i1 ? nG01 : nG02);
            const double g1 = i0 ? nG01 : (i1 ? nG11 : nG12);
            const double g2 = i0 ? nG02 : (i1 ? nG12 : nG22);
            double Dt0, Dt1, Dt2, G0, G1, G2, Nd, No;
            if (!fast) {
                const double J6 = 6 * J, J2 = 2 * J;
                Dt0 = d0 / J6; Dt1 = d1 / J6; Dt2 = d2 / J6;
                G0 = g0 / J2; G1 = g1 / J2; G2 = g2 / J2;
                Nd = J / 12; No = J / 24;                                         // shape_funcs.py:410-411
            } else {
                const double i6 = rJ * (1.0 / 6.0), i2 = 0.5 * rJ;
                Dt0 = d0 * i6; Dt1 = d1 * i6; Dt2 = d2 * i6;
                G0 = g0 * i2; G1 = g1 * i2; G2 = g2 * i2;
                Nd = J * (1.0 / 12.0); No = J * (1.0 / 24.0);
            }
            const double N0 = i0 ? Nd : No, N1 = i1 ? Nd : No, N2 = (i == 2) ? Nd : No;
            if (mode != 0) {
                double (&Mv)[6] = mode == 1 ? Bv : Av;
                Mv[0] = -c * Dt0; Mv[1] = -c * Dt1; Mv[2] = -c * Dt2;
                Mv[3] = c * (G0 + sigma * N0); Mv[4] = c * (G1 + sigma * N1); Mv[5] = c * (G2 + sigma * N2);
                if (mode == 1) return;
            }
            Bv[0] = Dt0; Bv[1] = Dt1; Bv[2] = Dt2;                                // fe_solver.py:209
            Bv[3] = __dsub_rn(G0, __dmul_rn(c, N0));                              // fe_solver.py:199
            Bv[4] = __dsub_rn(G1, __dmul_rn(c, N1));
            Bv[5] = __dsub_rn(G2, __dmul_rn(c, N2));
        }
    }
};

// One thread per matrix row.  The raw inputs of the NEXT incident element are requested before
// the arithmetic of the current one starts (software prefetch: the incidence -> connectivity ->
// coordinates chain is three dependent loads deep), and the row's slots live in a shared-memory
// tile that is written back with coalesced stores.
// SMEM = true: every 128-row block of the launch fits the tile (the normal case) and the
// accumulators are addressed in the shared window (LDS/STS, not generic LD/ST); SMEM = false is
// the fallback for unusually dense blocks, which accumulate straight into the (zeroed) output.
// MINB = resident CTAs per SM the register allocation is sized for.
// FIRST = true: bit 15 of a position marks the first contribution to its slot (lowest element of
// the row that touches it): it is stored, later ones are added -- no zero-fill pass, and 40-55 % of
// the shared-memory loads of the accumulate disappear (9 of the 12 contributions to an edge row
// are the only ones their slot ever receives).
template <class Op, bool WRITE_A, bool PREFETCH, bool SMEM, int MINB, bool FIRST>
__global__ void __launch_bounds__(ROWS_PER_CTA, MINB)
k_assemble_rows(RowCtx cx, Op op, double* __restrict__ A, double* __restrict__ B) {
    extern __shared__ double tile[];
    const int64_t r0 = (cx.first_block + blockIdx.x) * ROWS_PER_CTA;
    const int64_t r1 = min(r0 + ROWS_PER_CTA, cx.N);
    const int s0 = cx.rowptr[r0], s1 = cx.rowptr[r1];
    const int S = s1 - s0;
    double* tB;
    double* tA;
    if (SMEM) {
        tB = tile;
        tA = tile + cx.cap;
    } else {
        tB = B + s0;
        tA = WRITE_A ? A + s0 : nullptr;
    }
    if (!FIRST) {
        for (int t = threadIdx.x; t < S; t += ROWS_PER_CTA) {
            tB[t] = 0.0;
            if (WRITE_A) tA[t] = 0.0;
        }
        __syncthreads();   // (global fallback: the zero-fill above is visible CTA-wide after the barrier)
    }
    const int64_t r = r0 + threadIdx.x;
    const bool batch_rmw = cx.batch != 0;
    if (r < r1) {
        const int base = cx.rowptr[r] - s0;
        const int i0 = cx.incptr[r], i1 = cx.incptr[r + 1];
        if (i0 < i1) {
            uint32_t u = cx.inc[i0];
            RawTri cur = op.load(u >> 3);
            for (int i = i0; i < i1; ++i) {
                uint32_t un = u;
                RawTri nxt = cur;
                if (PREFETCH) {
                    un = (i + 1 < i1) ? cx.inc[i + 1] : u;
                    nxt = op.load(un >> 3);
                }
                const uint32_t* pp = reinterpret_cast<const uint32_t*>(cx.pos + (int64_t)i * 6);
                const uint32_t pk0 = pp[0], pk1 = pp[1], pk2 = pp[2];
                double Av[6], Bv[6];
                int cb[6];
                op.row(cur, (int)(u & 7u), Av, Bv, cb);
                if (FIRST && batch_rmw) {
                    // the six targets of one incidence are distinct slots of the row (no element lists an unknown twice when
                    // the first-touch flags are valid): all loads first, then the stores -- one dependent shared-memory round
                    // trip per incidence instead of twelve
                    int qs[6];
                    double ob[6], oa[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const int h = cb[j] >> 1;
                        const uint32_t w = h == 0 ? pk0 : (h == 1 ? pk1 : pk2);
                        const uint32_t pq = (cb[j] & 1) ? (w >> 16) : (w & 0xffffu);
                        const int q = base + (int)(pq & 0x7fffu);
                        const bool fr = (pq & 0x8000u) != 0;
                        qs[j] = fr ? ~q : q;
                        ob[j] = fr ? 0.0 : tB[q];
                        if (WRITE_A) oa[j] = fr ? 0.0 : tA[q];
                    }
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const bool fr = qs[j] < 0;
                        const int q = fr ? ~qs[j] : qs[j];
                        tB[q] = fr ? Bv[j] : ob[j] + Bv[j];
                        if (WRITE_A) tA[q] = fr ? Av[j] : oa[j] + Av[j];
                    }
                } else
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int h = cb[j] >> 1;
                    const uint32_t w = h == 0 ? pk0 : (h == 1 ? pk1 : pk2);
                    const uint32_t pq = (cb[j] & 1) ? (w >> 16) : (w & 0xffffu);
                    if (FIRST) {
                        const int q = base + (int)(pq & 0x7fffu);
                        if (pq & 0x8000u) {
                            tB[q] = Bv[j];
                            if (WRITE_A) tA[q] = Av[j];
                        } else {
                            tB[q] += Bv[j];
                            if (WRITE_A) tA[q] += Av[j];
                        }
                    } else {
                        const int q = base + (int)(pq & 0x7fffu);
                        tB[q] += Bv[j];
                        if (WRITE_A) tA[q] += Av[j];
                    }
                }
                if (PREFETCH) {
                    cur = nxt;
                    u = un;
                } else if (i + 1 < i1) {
                    u = cx.inc[i + 1];
                    cur = op.load(u >> 3);
                }
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int t = threadIdx.x; t < S; t += ROWS_PER_CTA) {
            B[s0 + t] = tB[t];
            if (WRITE_A) A[s0 + t] = tA[t];
        }
    }
}

struct AsmTuning {
    int n_ranges, prefetch, minb;
};

static const AsmTuning& asm_tuning() {
    static AsmTuning t = [] {
        AsmTuning v;
        const char* e = getenv("WS_ASM_RANGES");
        v.n_ranges = e ? std::max(1, std::min(atoi(e), (int)ws_pattern::NRANGE)) : 1;
        const char* f = getenv("WS_ASM_PREFETCH");
        v.prefetch = f ? atoi(f) : 0;
        const char* m = getenv("WS_ASM_MINB");
        v.minb = m ? atoi(m) : 5;   // 96 registers, 5 CTAs per SM: 5 % faster than 4 (114 registers), 6 and 8 spill and are slower
        return v;
    }();
    return t;
}

template <class Op, bool WRITE_A>
struct AsmKernels {
    using Fn = void (*)(RowCtx, Op, double*, double*);
    template <bool FIRST>
    static Fn smem_f(int prefetch, int minb) {
        if (prefetch) {
            if (minb >= 8) return k_assemble_rows<Op, WRITE_A, true, true, 8, FIRST>;
            if (minb >= 6) return k_assemble_rows<Op, WRITE_A, true, true, 6, FIRST>;
            if (minb >= 5) return k_assemble_rows<Op, WRITE_A, true, true, 5, FIRST>;
            return k_assemble_rows<Op, WRITE_A, true, true, 4, FIRST>;
        }
        if (minb >= 8) return k_assemble_rows<Op, WRITE_A, false, true, 8, FIRST>;
        if (minb >= 6) return k_assemble_rows<Op, WRITE_A, false, true, 6, FIRST>;
        if (minb >= 5) return k_assemble_rows<Op, WRITE_A, false, true, 5, FIRST>;
        return k_assemble_rows<Op, WRITE_A, false, true, 4, FIRST>;
    }
    static Fn smem(int prefetch, int minb, bool first) { return first ? smem_f<true>(prefetch, minb) : smem_f<false>(prefetch, minb); }
    static Fn global() { return k_assemble_rows<Op, WRITE_A, false, false, 4, false>; }
};

// The row blocks are launched in up to 16 contiguous groups, each with a shared-memory tile
// sized for its own densest block: sparse stretches of the matrix (edge rows, mid-side rows)
// then run at a much higher occupancy than the dense ones (vertex / node rows).
template <class Op, bool WRITE_A>
static void launch_rows(const ws_pattern* p, const Op& op, double* A, double* B, cudaStream_t st) {
    if (p->N == 0 || p->nnz == 0) return;
    const int per_slot = 8 * (WRITE_A ? 2 : 1);
    const AsmTuning& tn = asm_tuning();
    static const bool no_first = getenv("WS_ASM_NO_FIRST") && atoi(getenv("WS_ASM_NO_FIRST"));
    auto kern = AsmKernels<Op, WRITE_A>::smem(tn.prefetch, tn.minb, p->first_touch && !no_first);
    auto kern_g = AsmKernels<Op, WRITE_A>::global();
    WS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_TILE_BYTES));
    const int nblocks = ceil_div(p->N, ROWS_PER_CTA);
    const int per_group = std::max(1, ceil_div(nblocks, ws_pattern::NRANGE));
    const int groups_per_launch = ws_pattern::NRANGE / tn.n_ranges;
    for (int g = 0; g * per_group < nblocks; g += groups_per_launch) {
        const int b0 = g * per_group, b1 = std::min(nblocks, (g + groups_per_launch) * per_group);
        int cap = 32;
        for (int gg = g; gg < g + groups_per_launch && gg < ws_pattern::NRANGE; ++gg) cap = std::max(cap, p->range_block_slots[gg]);
        cap = (cap + 31) & ~31;
        static const int batch = getenv("WS_ASM_BATCH") ? atoi(getenv("WS_ASM_BATCH")) : 1;   // 0.277 -> 0.265 ms (vector), 0.271 -> 0.263 ms (scalar P2)
        RowCtx cx{p->rowptr.get(), p->incptr.get(), p->inc.get(), p->pos.get(), p->N, b0, cap, batch};
        if ((size_t)cap * per_slot <= MAX_TILE_BYTES)
            kern<<<b1 - b0, ROWS_PER_CTA, (size_t)cap * per_slot, st>>>(cx, op, A, B);
        else
            kern_g<<<b1 - b0, ROWS_PER_CTA, 0, st>>>(cx, op, A, B);   // a block denser than the largest tile
        count_launch(1);
    }
    WS_CUDA(cudaGetLastError());
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
                          const double* elem_k2n2, double* A_vals, double* B_vals, int flags,
                          int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K3 assemble scalar P2");
    // WS_ASM_TRANSPOSE is a no-op here: P2 scalar element matrices are bit-symmetric
    WS_REQUIRE(p, "pattern == NULL");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && dofs && elem_k2n2 && A_vals && B_vals, "null pointer");
    DeviceGuard g(device);
    ScalarP2 op{(const double2*)xy, dofs, elem_k2n2, (flags & WS_ASM_FAST) ? 1 : 0};
    launch_rows<ScalarP2, true>(p, op, A_vals, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_scalar_p1(const ws_pattern* p, const double* xy, const int32_t* dofs,
                          const double* elem_k2n2, double* A_vals, double* B_vals, int device,
                          void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && dofs && elem_k2n2 && A_vals && B_vals, "null pointer");
    DeviceGuard g(device);
    ScalarP1 op{(const double2*)xy, dofs, elem_k2n2};
    launch_rows<ScalarP1, true>(p, op, A_vals, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_mass_p2(const ws_pattern* p, const double* xy, const int32_t* dofs, double* B_vals,
                        int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && dofs && B_vals, "null pointer");
    DeviceGuard g(device);
    ScalarP2 op{(const double2*)xy, dofs, nullptr, 0};
    launch_rows<ScalarP2, false>(p, op, nullptr, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_vector_p1(const ws_pattern* p, const double* xy, const int32_t* tris,
                          const double* elem_k2n2, int64_t Ntt, double* A_vals, double* B_vals,
                          int flags, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K4 assemble vector");
    WS_REQUIRE(p, "pattern == NULL");
    WS_REQUIRE(Ntt >= 0 && Ntt <= p->N, "0 <= Ntt <= N");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && tris && elem_k2n2 && A_vals && B_vals, "null pointer");
    DeviceGuard g(device);
    VectorP1 op{(const double2*)xy, tris, elem_k2n2, (flags & WS_ASM_TRANSPOSE) ? 1 : 0, 0, 0.0, (flags & WS_ASM_FAST) ? 1 : 0};
    launch_rows<VectorP1, true>(p, op, A_vals, B_vals, (cudaStream_t)stream);
    WS_API_END
}

int ws_assemble_vector_qd(const ws_pattern* p, const double* xy, const int32_t* tris,
                          const double* elem_k2n2, int64_t Ntt, double sigma, double* M_vals,
                          double* B_vals, int flags, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K4 assemble vector (quasi-definite)");
    WS_REQUIRE(p, "pattern == NULL");
    WS_REQUIRE(Ntt >= 0 && Ntt <= p->N, "0 <= Ntt <= N");
    if (p->nnz == 0) return WS_OK;
    WS_REQUIRE(xy && tris && elem_k2n2 && M_vals, "null pointer");
    DeviceGuard g(device);
    const int fast = (flags & WS_ASM_FAST) ? 1 : 0;
    if (B_vals) {
        VectorP1 op{(const double2*)xy, tris, elem_k2n2, 0, 2, sigma, fast};
        launch_rows<VectorP1, true>(p, op, M_vals, B_vals, (cudaStream_t)stream);
    } else {
        VectorP1 op{(const double2*)xy, tris, elem_k2n2, 0, 1, sigma, fast};
        launch_rows<VectorP1, false>(p, op, nullptr, M_vals, (cudaStream_t)stream);
    }
    WS_API_END
}

// x = T y (mode 0):  x_t[i] = y_t[i] - (y_z[hi] - y_z[lo]) / len_i ,  x_z = y_z
// y = T^T b (mode 1): y_t = b_t ,  y_z[j] = b_z[j] - sum_{i: hi=j} b_t[i]/len_i + sum_{i: lo=j} b_t[i]/len_i
// Edge lengths and the node -> incident-edge lists (ascending edge id, so the sum order is fixed)
// are built once per (edge table, coordinates) and cached in the pattern object: both modes are
// then two short gathers per entry instead of a walk over the pattern row with a square root per edge.
__global__ void k_vt_len(int64_t Ntt, const int2* __restrict__ edges, const double2* __restrict__ xy, double* __restrict__ len,
                         int32_t* __restrict__ deg) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Ntt) return;
    const int2 e = edges[i];
    const double2 a = xy[e.x], b = xy[e.y];
    len[i] = sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y));
    atomicAdd(&deg[e.x], 1);
    atomicAdd(&deg[e.y], 1);
}
__global__ void k_vt_fill(int64_t Ntt, const int2* __restrict__ edges, const int32_t* __restrict__ nptr, int32_t* __restrict__ cur,
                          int32_t* __restrict__ nadj) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Ntt) return;
    const int2 e = edges[i];
    nadj[nptr[e.x] + atomicAdd(&cur[e.x], 1)] = (int32_t)(i << 1);
    nadj[nptr[e.y] + atomicAdd(&cur[e.y], 1)] = (int32_t)(i << 1) | 1;
}
__global__ void k_vt_sort(int64_t Nzz, const int32_t* __restrict__ nptr, int32_t* __restrict__ nadj) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= Nzz) return;
    const int a = nptr[j], b = nptr[j + 1];
    for (int i = a + 1; i < b; ++i) {          // insertion sort of a handful of entries
        const int32_t v = nadj[i];
        int k = i - 1;
        while (k >= a && nadj[k] > v) { nadj[k + 1] = nadj[k]; --k; }
        nadj[k + 1] = v;
    }
}
__global__ void k_vector_T(int64_t N, int64_t Ntt, int mode, const int2* __restrict__ edges, const double* __restrict__ len,
                           const int32_t* __restrict__ nptr, const int32_t* __restrict__ nadj, const double* __restrict__ in,
                           double* __restrict__ out) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= N) return;
    double v = in[r];
    if (r < Ntt) {
        if (mode == 0) {
            const int2 e = edges[r];
            v -= (in[Ntt + e.y] - in[Ntt + e.x]) / len[r];
        }
    } else if (mode == 1) {
        const int64_t j = r - Ntt;
        for (int s = nptr[j]; s < nptr[j + 1]; ++s) {
            const int32_t a = nadj[s];
            const int32_t c = a >> 1;
            v += ((a & 1) ? -in[c] : in[c]) / len[c];
        }
    }
    out[r] = v;
}

static void vt_prepare(const ws_pattern* p, const int32_t* edges, const double* xy, int64_t Ntt, cudaStream_t st) {
    if (p->vt_edges == edges && p->vt_xy == xy && p->vt_Ntt == Ntt) return;
    const int64_t Nzz = p->N - Ntt;
    p->vt_len.alloc(std::max<int64_t>(Ntt, 1), st);
    p->vt_nptr.alloc(Nzz + 1, st);
    p->vt_nadj.alloc(std::max<int64_t>(2 * Ntt, 1), st);
    DevBuf<int32_t> deg(Nzz + 1, st), cur(Nzz + 1, st);
    WS_CUDA(cudaMemsetAsync(deg.get(), 0, deg.bytes(), st));
    WS_CUDA(cudaMemsetAsync(cur.get(), 0, cur.bytes(), st));
    const int T = 256;
    if (Ntt) k_vt_len<<<ceil_div(Ntt, T), T, 0, st>>>(Ntt, (const int2*)edges, (const double2*)xy, p->vt_len.get(), deg.get());
    size_t bytes = 0;
    WS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, deg.get(), p->vt_nptr.get(), Nzz + 1, st));
    DevBuf<char> tmp(bytes, st);
    WS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, deg.get(), p->vt_nptr.get(), Nzz + 1, st));
    if (Ntt) k_vt_fill<<<ceil_div(Ntt, T), T, 0, st>>>(Ntt, (const int2*)edges, p->vt_nptr.get(), cur.get(), p->vt_nadj.get());
    if (Nzz) k_vt_sort<<<ceil_div(Nzz, T), T, 0, st>>>(Nzz, p->vt_nptr.get(), p->vt_nadj.get());
    WS_CUDA(cudaGetLastError());
    count_launch(3);
    p->vt_edges = edges;
    p->vt_xy = xy;
    p->vt_Ntt = Ntt;
    // (deg / cur / tmp are released stream-ordered: cudaFreeAsync after the kernels that use them)
}

int ws_vector_T(const ws_pattern* p, const int32_t* edges, const double* xy, int64_t Ntt, int mode,
                const double* in, double* out, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p && edges && xy && in && out, "null pointer");
    WS_REQUIRE(in != out, "in-place application is not supported");
    WS_REQUIRE(Ntt >= 0 && Ntt <= p->N && (mode == 0 || mode == 1), "bad Ntt / mode");
    DeviceGuard g(device);
    if (p->N) {
        vt_prepare(p, edges, xy, Ntt, (cudaStream_t)stream);
        k_vector_T<<<ceil_div(p->N, 256), 256, 0, (cudaStream_t)stream>>>(p->N, Ntt, mode, (const int2*)edges, p->vt_len.get(),
                                                                          p->vt_nptr.get(), p->vt_nadj.get(), in, out);
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
