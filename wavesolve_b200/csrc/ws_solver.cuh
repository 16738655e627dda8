// Shared declarations of the sparse direct solver (K8): symbolic structure produced on the
// host, task records consumed by the generic tile kernels on the device.
//
// Replaces SuperLU behind scipy splu / spsolve (reference fe_solver.py:328 via eigsh sigma
// mode, fe_solver.py:398) -- cuDSS is not present on the target image, so the solver is ours:
// element-based geometric nested dissection, multifrontal LDL^T without pivoting (the scalar
// shift-invert matrix is definite up to a few modes; the vector matrix is made quasi-definite
// by the congruence described in ws_vector.cu), dense fp64 fronts, selective inversion of the
// pivot blocks so that both triangular sweeps are batched dense mat-vecs.
#pragma once
#include <cstdint>
#include <vector>

#include "ws_common.cuh"

namespace ws {

constexpr int NB = 32;      // block-column width of the panel factorisation
constexpr int GT = 64;      // GEMM output tile (GT x GT), one CTA per tile
constexpr int PANEL_ROWS = 128;
constexpr int GEMV_ROWS = 32;
constexpr int GEMVT_COLS = 8;

// ---- host-side symbolic structure -----------------------------------------------------------
struct Symbolic {
    int64_t N = 0, Ne = 0;
    int depth = 0;
    int64_t nfronts = 0;                 // heap ids 1 .. nfronts-1 are used (index 0 unused)
    std::vector<int32_t> perm;           // new -> old
    std::vector<int32_t> iperm;          // old -> new
    std::vector<int32_t> p, q;           // own / boundary size per front
    std::vector<int64_t> start;          // first own new-index
    std::vector<int64_t> bndoff;         // offset into bnd[] / rel[]
    std::vector<int32_t> bnd;            // boundary new-indices, ascending, per front
    std::vector<int32_t> rel;            // position of bnd entries inside the parent's front
    std::vector<int64_t> Loff, Uoff, woff;   // offsets (in doubles) into panel / update / work storage
    std::vector<int32_t> node_of_new;    // owner front of every new index
    int64_t Ltotal = 0, Utotal = 0, Wtotal = 0;
    int64_t nnzL = 0;
    double flops = 0;
    int maxfront = 0, maxp = 0;
    double t_kd = 0, t_order = 0, t_bnd = 0;
};

}  // namespace ws

// Host analysis (ws_symbolic_create): every array of S is filled on the host.
// Device analysis (ws_symbolic_create_device): the per-front arrays of S (p, q, start, bndoff,
// Loff, Uoff, woff) are on the host, the large index arrays live in the d_* buffers only.
struct ws_symbolic {
    ws::Symbolic S;
    bool on_device = false;
    int device = 0;
    int64_t sumq = 0;
    ws::DevBuf<int32_t> d_perm, d_iperm, d_node_of_new, d_bnd, d_rel;
};

namespace ws {

// ---- device task records ---------------------------------------------------------------------
// C(m x n) = beta*C + alpha * A(m x k) * op(B), all column-major.
//   NT: op(B)[kk][j] = B[j + kk*ldb] * (dscale ? dscale[kk] : 1)
//   NN: op(B)[kk][j] = B[kk + j*ldb]
struct GemmTask {
    const double* A;
    const double* B;
    double* C;
    const double* dscale;
    int32_t lda, ldb, ldc;
    int32_t m, n, k;
    int32_t flags;       // bit0: accumulate into C (beta = 1); bit1: alpha = -1; bit2: B is N (not T)
    int32_t pad;
};
enum { GF_ACCUM = 1, GF_NEG = 2, GF_BN = 4 };

// factor the NB x NB diagonal block at (k0,k0) of a panel (redundantly per task) and apply
// L_kk^-T D^-1 to rows [r0, r0+nr) of block column k0.
struct PanelTask {
    double* L;           // panel base (m x p, ld)
    double* d;           // global pivot array at start + k0
    int32_t ld, k0, nb, r0, nr, write_diag;
};

struct LevelRange {
    int64_t begin = 0, count = 0;
};

}  // namespace ws
