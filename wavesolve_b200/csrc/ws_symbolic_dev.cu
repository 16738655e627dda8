// K8, symbolic phase on the DEVICE: the same element-based geometric nested dissection as
// ws_symbolic.cu (which stays as the host model the CPU tests check), expressed as sorts, scans
// and per-item kernels so that a 1M-element mesh is analysed in a few milliseconds and the
// large index arrays (permutation, front boundaries, parent-relative positions) are born in HBM
// instead of being built by host threads and pushed over PCIe.
//
//  1. k-d bisection: `depth` rounds of  (bounding box per segment) -> (segment, quantised
//     coordinate along the longer axis) keys -> stable radix sort.  Segments are the halving
//     ranges [lo, hi) -> [lo, mid) , [mid, hi) of the position array, so the segment of a
//     position is a function of the position alone and the tree is perfectly balanced.
//  2. owner of an unknown = lowest common ancestor of the leaves of its elements (atomic
//     min / max of the leaf index), elimination order = stable sort by (level-major rank of the
//     owner, class).
//  3. boundary(t) = unknowns touched by elements of subtree(t) that are owned by proper
//     ancestors of t: every (element, unknown) incidence emits one (front, new index) key per
//     tree level between the element's leaf and the owner; sort + unique gives all boundary
//     lists, ascending, in one array.
//  4. parent-relative positions by binary search in the parent's list; panel / update / work
//     offsets on the host from the per-front sizes (2^(depth+1) numbers).
//
// Replaces COLAMD + the symbolic phase of SuperLU behind scipy splu (fe_solver.py:328, :398).
#include <algorithm>
#include <climits>
#include <cub/cub.cuh>

#include "ws_solver.cuh"

namespace ws {

namespace {

constexpr int KD_BITS = 20;     // quantisation of the split coordinate inside a segment's box

__device__ __forceinline__ uint64_t order_bits(double v) {
    uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double unorder_bits(uint64_t u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

// segment (node index inside its level) of position i after `level` halvings of [0, Ne)
__device__ __forceinline__ int32_t seg_of(int64_t Ne, int level, int64_t i) {
    int64_t lo = 0, hi = Ne;
    int32_t s = 0;
    for (int l = 0; l < level; ++l) {
        const int64_t mid = (lo + hi) / 2;
        if (i < mid) { hi = mid; s = 2 * s; }
        else { lo = mid; s = 2 * s + 1; }
    }
    return s;
}

__global__ void k_centroids(const int32_t* __restrict__ dofs, int64_t Ne, const double2* __restrict__ xy, int64_t Np,
                            int vert_col, int64_t vert_offset, double2* __restrict__ cen, int32_t* __restrict__ ids,
                            int* __restrict__ flags) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= Ne) return;
    double cx = 0.0, cy = 0.0;
    for (int a = 0; a < 3; ++a) {
        const int64_t v = (int64_t)dofs[e * 6 + vert_col + a] - vert_offset;
        if (v < 0 || v >= Np) { flags[0] = 1; continue; }
        const double2 P = xy[v];
        cx += P.x;
        cy += P.y;
    }
    cen[e] = make_double2(cx / 3.0, cy / 3.0);
    ids[e] = (int32_t)e;
}

__global__ void k_bbox_init(int64_t nseg, uint64_t* __restrict__ bb) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    bb[4 * s + 0] = ~0ull; bb[4 * s + 1] = 0ull; bb[4 * s + 2] = ~0ull; bb[4 * s + 3] = 0ull;
}

// bounding box of every segment: min/max of the order-preserving bit patterns.  A block (or
// warp) that lies inside one segment reduces first and issues four atomics.
__global__ void __launch_bounds__(256) k_bbox(int64_t Ne, int level, const int32_t* __restrict__ ids,
                                              const double2* __restrict__ cen, uint64_t* __restrict__ bb) {
    __shared__ uint64_t red[4][8];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool on = i < Ne;
    const int64_t ii = on ? i : Ne - 1;
    const int32_t s = seg_of(Ne, level, ii);
    const double2 c = cen[ids[ii]];
    uint64_t x0 = order_bits(c.x), x1 = x0, y0 = order_bits(c.y), y1 = y0;
    const int64_t b0 = blockIdx.x * (int64_t)blockDim.x, b1 = min(b0 + (int64_t)blockDim.x, Ne) - 1;
    const bool block_uniform = seg_of(Ne, level, b0) == seg_of(Ne, level, b1);
    const int32_t sw = __shfl_sync(0xffffffffu, s, 0);
    const bool warp_uniform = __all_sync(0xffffffffu, s == sw);
    if (block_uniform || warp_uniform) {
        for (int o = 16; o; o >>= 1) {
            x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
            x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
            y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
        }
    }
    if (block_uniform) {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { red[0][w] = x0; red[1][w] = x1; red[2][w] = y0; red[3][w] = y1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int nw = (blockDim.x + 31) >> 5;
            for (int j = 1; j < nw; ++j) {
                x0 = min(x0, red[0][j]); x1 = max(x1, red[1][j]);
                y0 = min(y0, red[2][j]); y1 = max(y1, red[3][j]);
            }
            atomicMin((unsigned long long*)&bb[4 * s + 0], (unsigned long long)x0);
            atomicMax((unsigned long long*)&bb[4 * s + 1], (unsigned long long)x1);
            atomicMin((unsigned long long*)&bb[4 * s + 2], (unsigned long long)y0);
            atomicMax((unsigned long long*)&bb[4 * s + 3], (unsigned long long)y1);
        }
    } else if (warp_uniform) {
        if ((threadIdx.x & 31) == 0) {
            atomicMin((unsigned long long*)&bb[4 * s + 0], (unsigned long long)x0);
            atomicMax((unsigned long long*)&bb[4 * s + 1], (unsigned long long)x1);
            atomicMin((unsigned long long*)&bb[4 * s + 2], (unsigned long long)y0);
            atomicMax((unsigned long long*)&bb[4 * s + 3], (unsigned long long)y1);
        }
    } else if (on) {
        atomicMin((unsigned long long*)&bb[4 * s + 0], (unsigned long long)x0);
        atomicMax((unsigned long long*)&bb[4 * s + 1], (unsigned long long)x1);
        atomicMin((unsigned long long*)&bb[4 * s + 2], (unsigned long long)y0);
        atomicMax((unsigned long long*)&bb[4 * s + 3], (unsigned long long)y1);
    }
}

__global__ void k_kd_keys(int64_t Ne, int level, const int32_t* __restrict__ ids, const double2* __restrict__ cen,
                          const uint64_t* __restrict__ bb, uint64_t* __restrict__ keys) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Ne) return;
    const int32_t s = seg_of(Ne, level, i);
    const double x0 = unorder_bits(bb[4 * s + 0]), x1 = unorder_bits(bb[4 * s + 1]);
    const double y0 = unorder_bits(bb[4 * s + 2]), y1 = unorder_bits(bb[4 * s + 3]);
    const double2 c = cen[ids[i]];
    const bool ax = (x1 - x0) >= (y1 - y0);          // split along the longer side of the box
    const double lo = ax ? x0 : y0, ext = ax ? x1 - x0 : y1 - y0, v = ax ? c.x : c.y;
    double u = ext > 0.0 ? (v - lo) / ext : 0.0;
    u = fmin(fmax(u, 0.0), 1.0);
    const uint64_t kq = (uint64_t)(u * (double)((1u << KD_BITS) - 1));
    keys[i] = ((uint64_t)(uint32_t)s << KD_BITS) | kq;
}

__global__ void k_leaf_of(int64_t Ne, int depth, const int32_t* __restrict__ ids, int32_t* __restrict__ leaf_of) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < Ne) leaf_of[ids[i]] = seg_of(Ne, depth, i);
}

__global__ void k_fill_i32(int64_t n, int32_t v, int32_t* __restrict__ a) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void k_owner_minmax(const int32_t* __restrict__ dofs, int64_t n_inc, int64_t N, const int32_t* __restrict__ leaf_of,
                               int32_t* __restrict__ lo, int32_t* __restrict__ hi, int* __restrict__ flags) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_inc) return;
    const int32_t d = dofs[i];
    if (d < 0 || d >= N) { flags[1] = 1; return; }
    const int32_t lf = leaf_of[i / 6];
    atomicMin(&lo[d], lf);
    atomicMax(&hi[d], lf);
}

// level-major, bottom-up rank of front nd (leaves first, root last)
__device__ __forceinline__ int64_t prank(int depth, int64_t nd) {
    const int l = 63 - __clzll((long long)nd);
    return (((int64_t)1 << (depth + 1)) - ((int64_t)1 << (l + 1))) + nd - ((int64_t)1 << l);
}
__device__ __forceinline__ int32_t node_of_rank(int depth, int64_t r) {
    for (int l = depth; l >= 0; --l) {
        const int64_t base = ((int64_t)1 << (depth + 1)) - ((int64_t)1 << (l + 1));
        if (r < base + ((int64_t)1 << l)) return (int32_t)(((int64_t)1 << l) + (r - base));
    }
    return 1;
}

__global__ void k_bucket(int64_t N, int depth, int64_t cls_start, const int32_t* __restrict__ lo, const int32_t* __restrict__ hi,
                         uint32_t* __restrict__ bucket, int32_t* __restrict__ iota, long long* __restrict__ missing) {
    const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= N) return;
    iota[d] = (int32_t)d;
    if (hi[d] < 0) {
        atomicMin(missing, (long long)d);
        bucket[d] = 0;
        return;
    }
    const uint32_t x = (uint32_t)(lo[d] ^ hi[d]);
    const int hb = x ? 32 - __clz(x) : 0;
    const int64_t nd = (((int64_t)1 << depth) + lo[d]) >> hb;
    bucket[d] = (uint32_t)(2 * prank(depth, nd) + (d >= cls_start ? 0 : 1));
}

__global__ void k_after_order(int64_t N, int depth, const uint32_t* __restrict__ sbucket, const int32_t* __restrict__ perm,
                              int32_t* __restrict__ iperm, int32_t* __restrict__ node_of_new) {
    const int64_t nw = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (nw >= N) return;
    iperm[perm[nw]] = (int32_t)nw;
    node_of_new[nw] = node_of_rank(depth, sbucket[nw] >> 1);
}

// start / p of every front from the sorted bucket ids (bucket 2r = class-0 unknowns of the front
// with rank r, bucket 2r+1 = its class-1 unknowns)
__global__ void k_front_own(int64_t nfronts, int depth, int64_t N, const uint32_t* __restrict__ sbucket,
                            int64_t* __restrict__ start, int32_t* __restrict__ p) {
    const int64_t nd = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (nd >= nfronts) return;
    if (nd == 0) { start[0] = 0; p[0] = 0; return; }
    const uint32_t b0 = (uint32_t)(2 * prank(depth, nd)), b1 = b0 + 2;
    int64_t lo = 0, hi = N;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sbucket[mid] < b0) lo = mid + 1; else hi = mid; }
    const int64_t s0 = lo;
    hi = N;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sbucket[mid] < b1) lo = mid + 1; else hi = mid; }
    start[nd] = s0;
    p[nd] = (int32_t)(lo - s0);
}

__global__ void k_count_pairs(const int32_t* __restrict__ dofs, int64_t n_inc, int depth, const int32_t* __restrict__ iperm,
                              const int32_t* __restrict__ node_of_new, int64_t* __restrict__ cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_inc) return;
    const int32_t o = node_of_new[iperm[dofs[i]]];
    cnt[i] = depth - (31 - __clz(o));
}

__global__ void k_emit_pairs(const int32_t* __restrict__ dofs, int64_t n_inc, int depth, const int32_t* __restrict__ iperm,
                             const int32_t* __restrict__ node_of_new, const int32_t* __restrict__ leaf_of,
                             const int64_t* __restrict__ off, uint64_t* __restrict__ keys) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_inc) return;
    const int32_t g = iperm[dofs[i]];
    const int n = depth - (31 - __clz(node_of_new[g]));
    uint64_t t = ((uint64_t)1 << depth) + (uint64_t)leaf_of[i / 6];
    uint64_t* out = keys + off[i];
    for (int j = 0; j < n; ++j, t >>= 1) out[j] = (t << 32) | (uint32_t)g;
}

__global__ void k_bnd_split(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ bnd) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) bnd[i] = (int32_t)(uint32_t)keys[i];
}

__global__ void k_front_bounds(int64_t nfronts, int64_t n, const uint64_t* __restrict__ keys, int64_t* __restrict__ bndoff) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > nfronts) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)(keys[mid] >> 32) < t) lo = mid + 1; else hi = mid; }
    bndoff[t] = lo;
}

__global__ void k_rel(int64_t n, const uint64_t* __restrict__ keys, const int64_t* __restrict__ bndoff,
                      const int64_t* __restrict__ start, const int32_t* __restrict__ p, const int32_t* __restrict__ bnd,
                      int32_t* __restrict__ rel, int* __restrict__ flags) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = (int64_t)(keys[i] >> 32);
    const int32_t g = (int32_t)(uint32_t)keys[i];
    if (c < 2) { rel[i] = -1; return; }
    const int64_t par = c >> 1;
    const int64_t ps = start[par], pe = ps + p[par];
    if (g < pe) {
        if (g < ps) flags[2] = 1;                 // would mean the unknown is owned below the parent
        rel[i] = (int32_t)(g - ps);
        return;
    }
    int64_t lo = bndoff[par], hi = bndoff[par + 1];
    const int64_t b0 = lo, b1 = hi;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (bnd[mid] < g) lo = mid + 1; else hi = mid; }
    if (lo >= b1 || bnd[lo] != g) flags[2] = 1;   // the parent's front must contain the child's boundary
    rel[i] = (int32_t)(p[par] + (lo - b0));
}

template <class K, class V>
void sort_pairs(DevBuf<K>& k0, DevBuf<K>& k1, DevBuf<V>& v0, DevBuf<V>& v1, int64_t n, int end_bit, DevBuf<char>& tmp,
                cudaStream_t st) {
    if (n <= 0) return;
    cub::DoubleBuffer<K> dk(k0.get(), k1.get());
    cub::DoubleBuffer<V> dv(v0.get(), v1.get());
    size_t bytes = 0;
    WS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, n, 0, end_bit, st));
    if (bytes > tmp.bytes()) tmp.alloc(bytes, st);
    WS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), bytes, dk, dv, n, 0, end_bit, st));
    if (dk.Current() != k0.get()) std::swap(k0, k1);
    if (dv.Current() != v0.get()) std::swap(v0, v1);
}

}  // namespace

// Fills the per-front host arrays of Y and the device arrays of the handle.
void analyse_device(ws_symbolic* H, const int32_t* dofs, int64_t Ne, int64_t N, const double* xy, int64_t Np,
                    int vert_col, int64_t vert_offset, int64_t cls_start, int leaf_elems, cudaStream_t st) {
    Symbolic& S = H->S;
    WS_REQUIRE(Ne > 0 && N > 0, "analyse: empty problem");
    WS_REQUIRE(leaf_elems >= 1, "leaf_elems >= 1");
    WS_REQUIRE(Ne * 6 < ((int64_t)1 << 31), "6 * Ne must fit int32");
    int depth = 0;
    while ((Ne >> depth) > leaf_elems && depth < 24) ++depth;
    S.N = N;
    S.Ne = Ne;
    S.depth = depth;
    const int64_t nfronts = (int64_t)1 << (depth + 1);
    S.nfronts = nfronts;
    const int T = 256;
    const int64_t n_inc = Ne * 6;

    DevBuf<int> flags(4, st);
    DevBuf<long long> missing(1, st);
    WS_CUDA(cudaMemsetAsync(flags.get(), 0, flags.bytes(), st));
    const long long none = LLONG_MAX;
    WS_CUDA(cudaMemcpyAsync(missing.get(), &none, sizeof none, cudaMemcpyHostToDevice, st));
    DevBuf<char> tmp;

    // ---- 1. k-d bisection
    DevBuf<int32_t> leaf_of(Ne, st);
    {
        DevBuf<double2> cen(Ne, st);
        DevBuf<int32_t> ids0(Ne, st), ids1(Ne, st);
        DevBuf<uint64_t> key0(Ne, st), key1(Ne, st);
        DevBuf<uint64_t> bb(4 * std::max<int64_t>(1, (int64_t)1 << std::max(depth - 1, 0)), st);
        k_centroids<<<ceil_div(Ne, T), T, 0, st>>>(dofs, Ne, (const double2*)xy, Np, vert_col, vert_offset, cen.get(),
                                                   ids0.get(), flags.get());
        count_launch(1);
        for (int l = 0; l < depth; ++l) {
            const int64_t nseg = (int64_t)1 << l;
            k_bbox_init<<<ceil_div(nseg, T), T, 0, st>>>(nseg, bb.get());
            k_bbox<<<ceil_div(Ne, T), T, 0, st>>>(Ne, l, ids0.get(), cen.get(), bb.get());
            k_kd_keys<<<ceil_div(Ne, T), T, 0, st>>>(Ne, l, ids0.get(), cen.get(), bb.get(), key0.get());
            sort_pairs(key0, key1, ids0, ids1, Ne, KD_BITS + l, tmp, st);
            count_launch(3);
        }
        k_leaf_of<<<ceil_div(Ne, T), T, 0, st>>>(Ne, depth, ids0.get(), leaf_of.get());
        count_launch(1);
    }

    // ---- 2. owners and elimination order
    H->d_perm.alloc(N, st);
    H->d_iperm.alloc(N, st);
    H->d_node_of_new.alloc(N, st);
    DevBuf<int64_t> d_start(nfronts, st);
    DevBuf<int32_t> d_p(nfronts, st);
    {
        DevBuf<int32_t> lo(N, st), hi(N, st), iota(N, st);
        DevBuf<uint32_t> b0(N, st), b1(N, st);
        k_fill_i32<<<ceil_div(N, T), T, 0, st>>>(N, INT_MAX, lo.get());
        k_fill_i32<<<ceil_div(N, T), T, 0, st>>>(N, -1, hi.get());
        k_owner_minmax<<<ceil_div(n_inc, T), T, 0, st>>>(dofs, n_inc, N, leaf_of.get(), lo.get(), hi.get(), flags.get());
        k_bucket<<<ceil_div(N, T), T, 0, st>>>(N, depth, cls_start, lo.get(), hi.get(), b0.get(), iota.get(), missing.get());
        sort_pairs(b0, b1, iota, H->d_perm, N, depth + 2, tmp, st);
        // (after the swap logic above the sorted values are in `iota`, the scratch in d_perm)
        std::swap(iota, H->d_perm);
        k_after_order<<<ceil_div(N, T), T, 0, st>>>(N, depth, b0.get(), H->d_perm.get(), H->d_iperm.get(),
                                                    H->d_node_of_new.get());
        k_front_own<<<ceil_div(nfronts, T), T, 0, st>>>(nfronts, depth, N, b0.get(), d_start.get(), d_p.get());
        count_launch(6);
    }
    int h_flags[4];
    long long h_missing;
    WS_CUDA(cudaMemcpyAsync(h_flags, flags.get(), sizeof h_flags, cudaMemcpyDeviceToHost, st));
    WS_CUDA(cudaMemcpyAsync(&h_missing, missing.get(), sizeof h_missing, cudaMemcpyDeviceToHost, st));

    // ---- 3. boundaries
    DevBuf<int64_t> cnt(n_inc + 1, st), off(n_inc + 1, st);
    k_count_pairs<<<ceil_div(n_inc, T), T, 0, st>>>(dofs, n_inc, depth, H->d_iperm.get(), H->d_node_of_new.get(), cnt.get());
    WS_CUDA(cudaMemsetAsync(cnt.get() + n_inc, 0, sizeof(int64_t), st));
    {
        size_t bytes = 0;
        WS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt.get(), off.get(), n_inc + 1, st));
        if (bytes > tmp.bytes()) tmp.alloc(bytes, st);
        WS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), bytes, cnt.get(), off.get(), n_inc + 1, st));
    }
    int64_t M = 0;
    WS_CUDA(cudaMemcpyAsync(&M, off.get() + n_inc, sizeof M, cudaMemcpyDeviceToHost, st));
    WS_CUDA(cudaStreamSynchronize(st));
    count_launch(2);
    WS_REQUIRE(!h_flags[0], "analyse: vertex id out of range");
    WS_REQUIRE(!h_flags[1], "analyse: unknown id out of range");
    if (h_missing != LLONG_MAX)
        throw Error(WS_ERR_NUMERIC, "matrix is singular: unknown " + std::to_string(h_missing) +
                                        " belongs to no element (empty row)");
    int64_t sumq = 0;
    DevBuf<int64_t> d_bndoff(nfronts + 1, st);
    {
        DevBuf<uint64_t> k0(std::max<int64_t>(M, 1), st), k1(std::max<int64_t>(M, 1), st);
        DevBuf<int64_t> nsel(1, st);
        if (M > 0) {
            k_emit_pairs<<<ceil_div(n_inc, T), T, 0, st>>>(dofs, n_inc, depth, H->d_iperm.get(), H->d_node_of_new.get(),
                                                           leaf_of.get(), off.get(), k0.get());
            cub::DoubleBuffer<uint64_t> dk(k0.get(), k1.get());
            size_t bytes = 0;
            WS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, dk, M, 0, 32 + depth + 1, st));
            if (bytes > tmp.bytes()) tmp.alloc(bytes, st);
            WS_CUDA(cub::DeviceRadixSort::SortKeys(tmp.get(), bytes, dk, M, 0, 32 + depth + 1, st));
            if (dk.Current() != k0.get()) std::swap(k0, k1);
            bytes = 0;
            WS_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, k0.get(), k1.get(), nsel.get(), M, st));
            if (bytes > tmp.bytes()) tmp.alloc(bytes, st);
            WS_CUDA(cub::DeviceSelect::Unique(tmp.get(), bytes, k0.get(), k1.get(), nsel.get(), M, st));
            WS_CUDA(cudaMemcpyAsync(&sumq, nsel.get(), sizeof sumq, cudaMemcpyDeviceToHost, st));
            WS_CUDA(cudaStreamSynchronize(st));
            count_launch(3);
        }
        WS_REQUIRE(sumq < ((int64_t)1 << 31), "boundary storage exceeds 32-bit indexing");
        H->sumq = sumq;
        H->d_bnd.alloc(std::max<int64_t>(sumq, 1), st);
        H->d_rel.alloc(std::max<int64_t>(sumq, 1), st);
        k_front_bounds<<<ceil_div(nfronts + 1, T), T, 0, st>>>(nfronts, sumq, k1.get(), d_bndoff.get());
        if (sumq > 0) {
            k_bnd_split<<<ceil_div(sumq, T), T, 0, st>>>(sumq, k1.get(), H->d_bnd.get());
            k_rel<<<ceil_div(sumq, T), T, 0, st>>>(sumq, k1.get(), d_bndoff.get(), d_start.get(), d_p.get(), H->d_bnd.get(),
                                                   H->d_rel.get(), flags.get());
        }
        count_launch(3);
        WS_CUDA(cudaMemcpyAsync(h_flags, flags.get(), sizeof h_flags, cudaMemcpyDeviceToHost, st));
        // ---- 4. per-front sizes to the host; offsets and statistics there
        S.p.assign(nfronts, 0);
        S.q.assign(nfronts, 0);
        S.start.assign(nfronts, 0);
        S.bndoff.assign(nfronts + 1, 0);
        WS_CUDA(cudaMemcpyAsync(S.p.data(), d_p.get(), nfronts * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaMemcpyAsync(S.start.data(), d_start.get(), nfronts * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaMemcpyAsync(S.bndoff.data(), d_bndoff.get(), (nfronts + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaStreamSynchronize(st));   // k0 / k1 die here
    }
    if (h_flags[2]) throw Error(WS_ERR_STATE, "internal: separator ownership inconsistent");
    S.Loff.assign(nfronts, 0);
    S.Uoff.assign(nfronts, 0);
    S.woff.assign(nfronts, 0);
    int64_t Lt = 0, Ut = 0, Wt = 0;
    S.nnzL = 0;
    S.flops = 0;
    S.maxfront = 0;
    S.maxp = 0;
    for (int64_t t = 1; t < nfronts; ++t) {
        S.q[t] = (int32_t)(S.bndoff[t + 1] - S.bndoff[t]);
        const int64_t p = S.p[t], q = S.q[t], m = p + q;
        S.Loff[t] = Lt; Lt += (m * p + 1) & ~(int64_t)1;
        S.Uoff[t] = Ut; Ut += (q * q + 1) & ~(int64_t)1;
        S.woff[t] = Wt; Wt += (m + 1) & ~(int64_t)1;
        S.nnzL += m * p;
        S.flops += (double)p * p * p / 3.0 + (double)p * p * q + (double)p * q * q;
        S.maxfront = std::max<int>(S.maxfront, (int)m);
        S.maxp = std::max<int>(S.maxp, (int)p);
    }
    S.Ltotal = Lt; S.Utotal = Ut; S.Wtotal = Wt;
    H->on_device = true;
}

}  // namespace ws

using namespace ws;

extern "C" int ws_symbolic_create_device(const int32_t* dofs, int64_t Ne, int64_t N, const double* xy, int64_t Np,
                                         int vert_col, int64_t vert_offset, int64_t first_class_start, int leaf_elems,
                                         ws_symbolic** out, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K8a symbolic analysis");
    WS_REQUIRE(out, "out == NULL");
    *out = nullptr;
    WS_REQUIRE(dofs && xy, "null pointer");
    WS_REQUIRE(vert_col == 0 || vert_col == 3, "vert_col must be 0 or 3");
    WS_REQUIRE(N < ((int64_t)1 << 31), "N must fit int32");
    DeviceGuard g(device);
    configure_mempool(device);
    auto* s = new ws_symbolic();
    s->device = device;
    try {
        analyse_device(s, dofs, Ne, N, xy, Np, vert_col, vert_offset, first_class_start, leaf_elems, (cudaStream_t)stream);
    } catch (...) {
        delete s;
        throw;
    }
    *out = s;
    WS_API_END
}
