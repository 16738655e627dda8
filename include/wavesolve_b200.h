/*
 * wavesolve_b200 -- C ABI of the B200-native FEM hot path (assembly + generalized sparse
 * eigensolve) that sits behind wavesolve's Python functions.
 *
 * wavesolve has no FFI of its own (it is pure Python); the seams this library replaces are
 * the reference *functions* (paths relative to /root/reference/src/wavesolve):
 *
 *   ws_edges_first_appearance    <- waveguide.py:10-43   get_unique_edges
 *   ws_pattern_create / export   <- fe_solver.py:38-40,67-68 (COO -> CSR, scipy coo_tocsr +
 *                                   sum_duplicates) and fe_solver.py:170-176,212-213 (lil -> csc)
 *   ws_assemble_scalar_p2        <- fe_solver.py:47-55 + shape_funcs.py:192-246
 *   ws_assemble_mass_p2          <- fe_solver.py:71-99   construct_B_order2
 *   ws_assemble_scalar_p1        <- fe_solver.py:103-129 construct_AB_order1 + shape_funcs.py:403-428
 *   ws_assemble_vector_p1        <- fe_solver.py:178-210 + shape_funcs.py:310-428
 *   ws_prune_zeros               <- the "exact zeros are not stored" behaviour of lil_matrix
 *                                   behind fe_solver.py:196-199, 212-213
 *   ws_spmv                      <- ARPACK reverse communication ido=2 (M_matvec) inside
 *                                   eigsh (fe_solver.py:328)
 *   ws_overlap / ws_bnormalize   <- `u.T @ B @ v` with B = construct_B(mesh) (fe_solver.py:71-101,
 *                                   notes eq. 38) and the B-normalisation eigsh applies to its output
 *   ws_solver_*                  <- SuperLU splu/gssv inside eigsh sigma-mode and spsolve
 *                                   (fe_solver.py:328, fe_solver.py:398)
 *   ws_eig_*                     <- eigsh(A, M=B, k, which='SA', sigma)   (fe_solver.py:328, :266) and
 *                                   eigs((A - sigma B)^-1 B, k, which='SR') (fe_solver.py:398-400)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (host);
 *   - the caller allocates every output array; opaque handles have explicit destroy;
 *   - every call takes the CUDA device ordinal and a stream (cudaStream_t passed as void*);
 *     work is enqueued on that stream; calls that must return a size to the host
 *     (h_ outputs) synchronise that stream once, all others are asynchronous;
 *   - return value: 0 ok, <0 argument error, >0 CUDA / numerical error; nothing throws across
 *     the ABI; ws_last_error() returns a thread-local message for the last failure;
 *   - connectivity is int32, values are fp64, CSR/CSC indices are int32 (as scipy produces
 *     for these sizes);
 *   - handles are not thread-safe; use one host thread (or process) per GPU.
 */
#ifndef WAVESOLVE_B200_H
#define WAVESOLVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WS_OK 0
#define WS_ERR_ARG (-1)
#define WS_ERR_STATE (-2)
#define WS_ERR_CUDA 1
#define WS_ERR_NUMERIC 2
#define WS_ERR_NOCONV 3

/* flags of the assembly entry points */
#define WS_ASM_TRANSPOSE 1 /* slot (i,j) receives element entry (j,i): CSC values of the symmetric pattern */
#define WS_ASM_FAST 2      /* one reciprocal per element row instead of IEEE divisions (<= 2 ulp off the
                              reference arithmetic); for the eigen-solve pipeline, not for exported matrices */

typedef struct ws_pattern ws_pattern; /* CSR pattern + row-incidence lists of one mesh   */
typedef struct ws_symbolic ws_symbolic; /* host: nested-dissection ordering + front tree     */
typedef struct ws_solver ws_solver;   /* multifrontal LDL^T of A - sigma B               */

int ws_version(void);
const char* ws_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
long long ws_launch_count(void);

/* ---- K1: unique edges numbered by first appearance (waveguide.py:10-43) -------------
 * tris      [Ne*3] vertex ids of the triangles in mesh order (cells[1].data)
 * edges     [3*Ne*2] out, first Ntt rows valid: sorted vertex pair of edge id
 * edge_idx  [Ne*3]  out: id of local edges (0,1),(1,2),(2,0) of every triangle
 * h_Ntt     host out: number of unique edges                                        */
int ws_edges_first_appearance(const int32_t* tris, int64_t Ne, int64_t Np, int32_t* edges,
                              int32_t* edge_idx, int64_t* h_Ntt, int device, void* stream);

/* ---- K2: sparsity pattern ------------------------------------------------------------
 * dofs [Ne*6]: the six global unknowns of every element, elements in the order the
 * reference assembles them (material-major, fe_solver.py:36 / :179).
 *   scalar P2 : the six node ids (v0 v1 v2 m01 m12 m20)
 *   vector P1 : edge ids e01 e12 e20, then Ntt + v0, Ntt + v1, Ntt + v2
 * The pattern is the structural one (union of the 6x6 couplings), rows and columns sorted,
 * int32 -- identical to what scipy builds at fe_solver.py:67-68 (explicit zeros kept).   */
int ws_pattern_create(const int32_t* dofs, int64_t Ne, int64_t N, ws_pattern** out, int device,
                      void* stream);
int ws_pattern_query(const ws_pattern* p, int64_t* h_N, int64_t* h_nnz, int64_t* h_Ne);
int ws_pattern_export(const ws_pattern* p, int32_t* indptr /*N+1*/, int32_t* indices /*nnz*/,
                      int device, void* stream);
/* device pointers owned by the pattern (valid until destroy) */
int ws_pattern_arrays(const ws_pattern* p, const int32_t** indptr, const int32_t** indices);
int ws_pattern_destroy(ws_pattern* p);

/* ---- K3: scalar P2 assembly (fe_solver.py:47-55, shape_funcs.py:192-246) -----------------
 * xy [Np*2] node coordinates, dofs as given to ws_pattern_create, elem_k2n2 [Ne] = k^2 n_e^2
 * (fe_solver.py:51).  A_vals/B_vals [nnz] in the pattern's slot order.
 * flags: WS_ASM_TRANSPOSE | WS_ASM_FAST.                                                 */
int ws_assemble_scalar_p2(const ws_pattern* p, const double* xy, const int32_t* dofs,
                          const double* elem_k2n2, double* A_vals, double* B_vals, int flags,
                          int device, void* stream);
/* mass matrix only, all elements (fe_solver.py:71-99) */
int ws_assemble_mass_p2(const ws_pattern* p, const double* xy, const int32_t* dofs, double* B_vals,
                        int device, void* stream);

/* scalar P1 (fe_solver.py:103-129, shape_funcs.py:403-428).  The pattern is built from six
 * unknowns per element: pass dofs = (v0 v1 v2 v0 v1 v2) per linear triangle, to
 * ws_pattern_create and here.  A_e = k2n2 * LNN - LdNdN, B_e = LNN.                          */
int ws_assemble_scalar_p1(const ws_pattern* p, const double* xy, const int32_t* dofs,
                          const double* elem_k2n2, double* A_vals, double* B_vals, int device,
                          void* stream);

/* ---- K4: vector assembly, Whitney edges + P1 nodes (fe_solver.py:178-210) -----------------
 * tris [Ne*3] vertex ids (element order = dofs order), Ntt = number of edge unknowns.
 * A_vals/B_vals [nnz] on the pattern of B; A is non-zero only in the edge-edge block.  */
int ws_assemble_vector_p1(const ws_pattern* p, const double* xy, const int32_t* tris,
                          const double* elem_k2n2, int64_t Ntt, double* A_vals, double* B_vals,
                          int flags, int device, void* stream);

/* ---- zero pruning (lil_matrix semantics behind fe_solver.py:196-199, 212-213) -------------
 * Compacts (pattern, vals) to the entries with vals != 0.0.  out_* sized for nnz.        */
int ws_prune_zeros(const ws_pattern* p, const double* vals, int32_t* out_indptr /*N+1*/,
                   int32_t* out_indices, double* out_vals, int64_t* h_nnz_out, int device,
                   void* stream);

/* ---- K11: A(k) = k2scale * Mn2 - K on a shared pattern; y = alpha*x + beta*y on values --- */
int ws_axpby(int64_t n, double alpha, const double* x, double beta, double* y, int device,
             void* stream);

/* ---- K5: y = CSR(pattern, vals) * x --------------------------------------------------- */
int ws_spmv(const ws_pattern* p, const double* vals, const double* x, double* y, int device,
            void* stream);

/* ---- K8a: symbolic analysis of the sparse factorisation (host model; the device version follows) --
 * Replaces the COLAMD ordering + symbolic phase of SuperLU inside scipy splu / spsolve
 * (fe_solver.py:328 through eigsh sigma mode, fe_solver.py:398).  Element-based geometric
 * nested dissection: h_dofs [Ne*6] as for ws_pattern_create, h_xy [Np*2] node coordinates
 * (both HOST arrays); the element centroids are formed from the three vertex ids stored at
 * h_dofs[e*6 + vert_col + 0..2] - vert_offset (scalar P2: vert_col 0, offset 0; vector:
 * vert_col 3, offset Ntt).  Unknowns with id >= first_class_start are eliminated first
 * inside every front (vector problem: pass Ntt so nodes precede edges; scalar: pass 0).
 * The result depends on the mesh only and is reusable across wavelengths and shifts.
 * ws_symbolic_query: h_info[8] = N, depth, nfronts, nnz(L), sum q^2, max front, max pivot
 * block, sum q;  h_dinfo[4] = factor flops, seconds k-d, seconds ordering, seconds boundaries.
 * ws_symbolic_export: perm [N] (new -> old), p,q,start [nfronts] (heap ids, 0 unused),
 * bndoff [nfronts+1], bnd / rel [sum q]. Any output pointer may be NULL.                  */
int ws_symbolic_create(const int32_t* h_dofs, int64_t Ne, int64_t N, const double* h_xy, int64_t Np,
                       int vert_col, int64_t vert_offset, int64_t first_class_start, int leaf_elems,
                       ws_symbolic** out);
/* The same analysis on the DEVICE (the product path): dofs [Ne*6] and xy [Np*2] are device
 * arrays; k-d bisection by per-level radix sorts, ownership by atomic min/max, boundaries by
 * sort + unique of (front, unknown) keys.  The large index arrays stay in HBM and are handed to
 * ws_solver_create without crossing PCIe; query / export work as for the host analysis.  The
 * tree is the same balanced halving tree; ties at a split are broken by quantised coordinate
 * and stable order, so individual leaves may differ from the host analysis of the same mesh. */
int ws_symbolic_create_device(const int32_t* dofs, int64_t Ne, int64_t N, const double* xy, int64_t Np,
                              int vert_col, int64_t vert_offset, int64_t first_class_start, int leaf_elems,
                              ws_symbolic** out, int device, void* stream);
int ws_symbolic_query(const ws_symbolic* s, int64_t* h_info, double* h_dinfo);
int ws_symbolic_export(const ws_symbolic* s, int32_t* h_perm, int32_t* h_p, int32_t* h_q,
                       int64_t* h_start, int64_t* h_bndoff, int32_t* h_bnd, int32_t* h_rel);
int ws_symbolic_destroy(ws_symbolic* s);

/* ---- K8b: numeric factorisation and solves (device) ----------------------------------------
 * ws_solver_create : binds a pattern to a symbolic analysis, allocates all fronts (they stay
 *                    resident; ~6 kB per unknown at 1M elements) and uploads the task plan.
 * ws_solver_factor : M = A_vals - sigma*B_vals (values on the pattern; B_vals may be NULL for
 *                    M = A_vals), multifrontal LDL^T with STATIC pivoting (|pivot| < 1e-10 max|M_ij| is
 *                    replaced by +-that) + selective inversion, followed by an accuracy probe: one system
 *                    M x = b is solved with the new factors and its normwise backward error decides how many
 *                    steps of iterative refinement every later ws_solver_solve performs (0 when the factors
 *                    pass as they are: definite / quasi-definite matrices at the default shift).  This is what
 *                    SuperLU's partial pivoting buys the reference for interior shifts (target_neff,
 *                    fe_solver.py:313-316, :375-378).  Synchronises the stream to read the statistics; returns
 *                    WS_ERR_NUMERIC on non-finite pivots or when refinement cannot reach a backward error of
 *                    1e-11 (never a silently inaccurate factorisation).  While refine steps > 0 the value
 *                    arrays must stay alive until the last solve (the residual b - M x is formed from them).
 * ws_solver_quality: h_info[4] = perturbed pivots, refinement steps per solve, refinement steps the probe
 *                    needed, unique id of this solver object;  h_dinfo[4] = probe backward error before / after refinement, max|M_ij|,
 *                    pivot threshold.
 * ws_solver_solve  : x = M^-1 b, nrhs right-hand sides stored contiguously [nrhs*N]; asynchronous.
 *                    Right-hand sides are swept in pairs that share every load of the inverted
 *                    panels (1.26 ms for two vs 1.03 ms for one at 1M triangles).
 * ws_solver_stats  : h_info[8] = nnz(L), fronts, max front, bytes, negative pivots, positive
 *                    pivots, bad pivots, tree depth; h_dinfo[3] = factor flops, min |pivot|,
 *                    max |pivot|.  (negative / positive pivots = inertia of M: Sylvester's law
 *                    gives the number of eigenvalues of (A,B) above sigma.)                    */
int ws_solver_create(const ws_pattern* p, const ws_symbolic* sym, ws_solver** out, int device,
                     void* stream);
int ws_solver_factor(ws_solver* s, const double* A_vals, const double* B_vals, double sigma,
                     int device, void* stream);
int ws_solver_solve(ws_solver* s, const double* b, double* x, int nrhs, int device, void* stream);
/* ws_solver_check : synchronises the stream and returns WS_ERR_STATE if a flag-synchronised sweep
 *                    gave up waiting for a front since the last solve (internal error guard).    */
int ws_solver_check(const ws_solver* s, int device, void* stream);
int ws_solver_stats(const ws_solver* s, int64_t* h_info, double* h_dinfo);
int ws_solver_quality(const ws_solver* s, int64_t* h_info, double* h_dinfo);
/* ws_solver_set_overlap: 1 (default) chains the sweep kernels of a solve with programmatic dependent launch and
 *                    per-front completion flags (a level overlaps the tail of the previous one: best for ONE large
 *                    problem); 0 uses plain stream order (best when several problems share the GPU on different
 *                    streams: early-resident, spinning CTAs only take slots from the other lanes).  Set before the
 *                    first solve that a ws_eig_workspace captures.                                               */
int ws_solver_set_overlap(ws_solver* s, int on);
int ws_solver_destroy(ws_solver* s);

/* ---- vector problem: quasi-definite form ---------------------------------------------------
 * The vector shift-invert matrix M = A - sigma*B is indefinite with singular pivot blocks
 * (curl-curl null space).  With G the discrete gradient (edge i = (lo,hi): G[i,hi] = 1/len_i,
 * G[i,lo] = -1/len_i) and T = [[I,-G],[0,I]], M' = T^T M T is symmetric quasi-definite:
 *     M' = [[ (k2n2-sigma) NeNe - curlcurl , -k2n2 NedN ], [ . , k2n2 (dNdN + sigma NN) ]]
 * so LDL^T needs no pivoting, and M^-1 = T M'^-1 T^T.
 * ws_assemble_vector_qd writes M' on the pattern (and, when B_vals != NULL, B in the same pass);
 * ws_vector_T applies T (mode 0) or T^T (mode 1).
 * edges [Ntt*2] as produced by ws_edges_first_appearance.                                    */
int ws_assemble_vector_qd(const ws_pattern* p, const double* xy, const int32_t* tris,
                          const double* elem_k2n2, int64_t Ntt, double sigma, double* M_vals,
                          double* B_vals, int flags, int device, void* stream);
int ws_vector_T(const ws_pattern* p, const int32_t* edges, const double* xy, int64_t Ntt, int mode,
                const double* in, double* out, int device, void* stream);

/* ---- K5-K7, K9, K10: Krylov eigensolver on OP = (A - sigma B)^-1 B ----------------------------
 * kind 0: B-orthogonal Arnoldi/Lanczos with thick restart for the symmetric-definite pencil
 *         (replaces eigsh(A, M=B, k=nev, which='SA', sigma), fe_solver.py:328); `s` holds the
 *         factorisation of A - sigma B; eigenvectors come back B-orthonormal.
 * kind 1: Euclidean Arnoldi with thick restart on the non-symmetric OP (replaces
 *         eigs(spsolve(A - sigma B, B), k, which='SR'), fe_solver.py:398-400); `s` holds the
 *         factorisation of the quasi-definite M' (ws_assemble_vector_qd) and edges/xy/Ntt
 *         describe T; eigenvectors come back with unit 2-norm.
 * kind 2: as kind 1 with block size 2: two basis vectors go through the operator per step and share
 *         one pair of sweeps (ws_solver_solve with nrhs = 2 costs 1.22x one right-hand side).  On the
 *         1M-triangle problem it needs 102-122 operator applications instead of 80, so it is not
 *         faster end to end (measured); kept as an option for operators whose application dominates more.
 * ncv <= 0 -> max(2 nev + 4, 20) (scipy: 2 nev + 1; three more vectors save 1-2 restart cycles at
 * tol = eps on these problems); tol <= 0 -> machine epsilon (scipy's default).
 * w [nev] (device): lambda = sigma + 1/theta, descending.  V [nev*N] (device): row j =
 * eigenvector j.  h_info[5] = converged, OP applications, restarts, complex Ritz values among
 * the wanted ones, kept directions at the last restart.  h_resid [nev]: Ritz residual estimates.
 * Returns WS_ERR_NOCONV (results still written) if fewer than nev pairs converged.
 * ws_host_eig_general: the small dense eigen-solver used for the projected problem (host;
 * exposed for testing).  V is row-major, column j = eigenvector j (complex pair: re, im).    */
int ws_eig_solve(const ws_pattern* p, ws_solver* s, const double* B_vals, double sigma, int kind,
                 int nev, int ncv, double tol, int maxit, uint64_t seed, const int32_t* edges,
                 const double* xy, int64_t Ntt, double* w, double* V, int64_t* h_info,
                 double* h_resid, int device, void* stream);
int ws_host_eig_general(int n, const double* h_A, double* h_wr, double* h_wi, double* h_V,
                        int symmetric);

/* ---- K11/K12 throughput path: several eigenproblems in lockstep on one GPU ----------------------------
 * A ~50k-triangle problem (lantern-demo.ipynb cell 4: 196 of them, run one after the other by the
 * reference) is a chain of small dependent kernels that fills a fraction of a B200.  ws_eig_solve_batch
 * drives `njobs` independent eigenproblems -- each with its own factorised solver, value arrays, outputs
 * and STREAM -- from the calling host thread, interleaving their restart cycles so that their kernels
 * overlap on the device (kind 0 / 1 only; arguments as for ws_eig_solve).  A failing job reports its own
 * rc / error text and does not disturb the others (rc == WS_ERR_NOCONV: w / V hold what the iteration had).
 * ws_solver_factor_begin / _end: the two halves of ws_solver_factor (enqueue everything incl. the accuracy
 * probe / read the statistics back, refine if needed) so that several factorisations can be in flight.
 * ws_eig_workspace: optional per-lane object that keeps the Krylov basis, the pinned staging buffers and the captured
 * CUDA graphs of the Arnoldi steps between the problems of one mesh (same pattern / solver / value array / sizes):
 * the ~ncv graph captures are then paid once per mesh, not once per wavelength.  The streams of the jobs that use
 * a workspace must be idle before it is destroyed.                                                              */
typedef struct ws_eig_workspace ws_eig_workspace;
int ws_eig_workspace_create(ws_eig_workspace** out);
int ws_eig_workspace_destroy(ws_eig_workspace* w);
typedef struct ws_eig_job {
    const ws_pattern* pattern;
    ws_solver* solver;
    const double* B_vals;
    double sigma;
    int kind, nev, ncv, maxit;
    double tol;
    uint64_t seed;
    const int32_t* edges;
    const double* xy;
    int64_t Ntt;
    double* w;      /* [nev]   device */
    double* V;      /* [nev*N] device */
    void* stream;   /* cudaStream_t of this job */
    double* resid;  /* [nev] host, may be NULL */
    ws_eig_workspace* workspace; /* may be NULL */
    int64_t info[5];
    int rc;
    char error[228];
} ws_eig_job;
int ws_eig_solve_batch(ws_eig_job* jobs, int njobs, int device);
int ws_solver_factor_begin(ws_solver* s, const double* A_vals, const double* B_vals, double sigma, int device,
                           void* stream);
int ws_solver_factor_end(ws_solver* s, int device, void* stream);

/* ---- K13-K15: resampling of modes onto points (grid / mesh-to-mesh interpolation) -------------
 * Replaces get_tri_idxs_KDtree / find_triangle_KDtree (fe_solver.py:848-872: one KD-tree query per
 * try and per point), get_tri_idxs / find_triangle (fe_solver.py:615-668), get_interp_weights
 * (fe_solver.py:670-696), get_interp_weights_vec (fe_solver.py:698-717) and the gather + sum of
 * interpolate / interpolate_vec (fe_solver.py:719-774).
 * ws_locator_create : uniform bin grid over the bounding box [bx0,bx1] x [by0,by1] of the mesh;
 *                     conn [Ne*stride] int32, the first three columns are the vertex ids.
 * ws_locate         : points are the grid xa[nx] x ya[ny] (ny > 0; point (i,j) at i*ny + j) or the
 *                     list (xa[k], ya[k]), k < nx (ny == 0).  tri_idx[...] = index of the containing
 *                     triangle under the reference's inclusive inside-test (fe_solver.py:597-612),
 *                     -1 if none or if maxr > 0 and |point| > maxr; ties between containing
 *                     triangles (points on edges / vertices): tie_rule 0 = nearest centroid (the
 *                     KD-tree search order), 1 = lowest index (the brute-force search order).
 * ws_interp_weights : kind 2: P2 weights [npts*6]; 1: P1 [npts*3]; 3: Whitney edge [npts*3*2];
 *                     NaN where tri_idx == -1.
 * ws_interp_apply   : out[pt, c] = sum_k field[dof[tri, k]] * weights[pt, k, c]; nd = 3 | 6 unknowns
 *                     per element, wdim = 1 | 2 components per weight, ncomp = 1 (real field) | 2
 *                     (complex128 field, interleaved).                                            */
typedef struct ws_locator ws_locator;
int ws_locator_create(const double* xy, const int32_t* conn, int64_t Ne, int stride, double bx0, double by0,
                      double bx1, double by1, ws_locator** out, int device, void* stream);
int ws_locator_destroy(ws_locator* L);
int ws_locate(const ws_locator* L, const double* xy, const int32_t* conn, int stride, const double* xa,
              const double* ya, int64_t nx, int64_t ny, double maxr, int tie_rule, int64_t* tri_idx, int device,
              void* stream);
int ws_interp_weights(const double* xy, const int32_t* conn, int64_t Ne, int stride, int kind, const double* xa,
                      const double* ya, int64_t nx, int64_t ny, const int64_t* tri_idx, double* weights, int device,
                      void* stream);
int ws_interp_apply(const int64_t* tri_idx, int64_t npts, int64_t Ne, const int32_t* dof, int stride, int nd, int wdim,
                    const double* weights, const double* field, int ncomp, double* out, int device, void* stream);

/* ---- K16/K17: overlaps in the mass-matrix inner product (second widening step) ---------------
 * The reference gives the user construct_B (fe_solver.py:71-101) and leaves `u.T @ B @ v` (notes eq. 38)
 * to scipy, one pair of host vectors at a time.  Modes are rows: U [nu x ldu], V [nv x ldv] (device,
 * ldu, ldv >= N); vals are values on the pattern (B from ws_assemble_mass_p2 / ws_assemble_*).
 * ws_overlap    : G[i*nv + j] = U_i^T B V_j   (device, row-major nu x nv; B V_j is read once per
 *                 pair of columns, the reduction order is fixed -> bitwise reproducible)
 * ws_bnormalize : d[i] = U_i^T B U_i (device); scale != 0 also divides row i by sqrt(d[i]) where
 *                 d[i] > 0 (rows with d <= 0 or NaN are left as they are).                          */
int ws_overlap(const ws_pattern* p, const double* vals, const double* U, int64_t ldu, int nu, const double* V,
               int64_t ldv, int nv, double* G, int device, void* stream);
int ws_bnormalize(const ws_pattern* p, const double* vals, double* U, int64_t ldu, int nu, double* d, int scale,
                  int device, void* stream);

/* ---- K12, host side: device -> pageable host memory for GB-sized gather payloads --------------------------
 * Two reusable page-locked staging blocks (64 MB each, kept for the life of the process); the DMA of chunk i+1
 * overlaps `nthreads` host threads copying chunk i into h_dst (<= 0: a default from the core count).  Synchronous:
 * returns when h_dst is complete.  (Page-locking a 2 GB landing buffer costs ~2 s; this path moves it in ~0.25 s.) */
int ws_copy_to_host(const void* d_src, void* h_dst, int64_t nbytes, int nthreads, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif
