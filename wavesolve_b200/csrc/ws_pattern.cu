// K1 (first-appearance edge numbering), K2 (CSR pattern + row-incidence lists), zero
// pruning.  All device-side; CUB radix sort / scan / select are the only building blocks
// that are not written here.
//
// Reference behaviour replaced:
//   waveguide.py:10-43     get_unique_edges  (O(E^2) python list scan)
//   fe_solver.py:38-40,67-68   COO triplets -> scipy coo_tocsr + sum_duplicates
//   fe_solver.py:170-176,212-213   lil_matrix accumulation -> tocsc (exact zeros dropped)
#include <algorithm>
#include <cub/cub.cuh>

#include "ws_common.cuh"

namespace ws {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void configure_mempool(int device) {
    static std::atomic<bool> done[64];
    if (device < 0 || device >= 64 || done[device].load()) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        // never make one stream wait for another just to recycle a freed block (concurrent sweep lanes
        // allocate and free solver storage on their own streams): take fresh memory instead
        int off = 0;
        cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
    }
    done[device].store(true);
}

// ---------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------
__global__ void k_inc_keys(const int32_t* __restrict__ dofs, int64_t n_inc, uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_inc) return;
    uint32_t e = (uint32_t)(i / 6), a = (uint32_t)(i % 6);
    keys[i] = ((uint64_t)(uint32_t)dofs[i] << 32) | (uint64_t)((e << 3) | a);
}

__global__ void k_pair_keys(const int32_t* __restrict__ dofs, int64_t Ne, uint64_t* __restrict__ keys) {
    // one thread per (element, a): writes the 6 keys (row = dof a, col = dof b)
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= Ne * 6) return;
    int64_t e = i / 6;
    const int32_t* d = dofs + e * 6;
    uint64_t row = (uint64_t)(uint32_t)dofs[i] << 32;
    uint64_t* out = keys + i * 6;
#pragma unroll
    for (int b = 0; b < 6; ++b) out[b] = row | (uint32_t)d[b];
}

__global__ void k_split_inc(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ inc) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) inc[i] = (uint32_t)keys[i];
}

// ptr[r] = first index i with (keys[i] >> 32) >= r, for r in [0, N]
__global__ void k_row_bounds(const uint64_t* __restrict__ keys, int64_t n, int64_t N, int32_t* __restrict__ ptr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > N) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> 32) < r) lo = mid + 1;
        else hi = mid;
    }
    ptr[r] = (int32_t)lo;
}

__global__ void k_low32(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)(uint32_t)keys[i];
}

__global__ void k_max_rowlen(const int32_t* __restrict__ rowptr, int64_t N, int* __restrict__ out) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int len = (r < N) ? rowptr[r + 1] - rowptr[r] : 0;
    for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len) atomicMax(out, len);
}

// out[0] = max entries of a `rows`-row block; out[1 + g] = the same over the blocks of group g
// (blocks are split into `ngroups` contiguous groups of `per_group` blocks)
__global__ void k_max_blockslots(const int32_t* __restrict__ rowptr, int64_t N, int rows, int per_group,
                                 int* __restrict__ out) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t r0 = b * rows;
    if (r0 >= N) return;
    int64_t r1 = min(r0 + rows, N);
    const int c = rowptr[r1] - rowptr[r0];
    atomicMax(out, c);
    atomicMax(out + 1 + (int)(b / per_group), c);
}

// pos[6*i+b] = position of column dofs[e*6+b] inside row r (r = row of incidence i)
__global__ void k_positions(const uint64_t* __restrict__ inc_keys, int64_t n_inc,
                            const int32_t* __restrict__ dofs, const int32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ colidx, uint16_t* __restrict__ pos) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_inc * 6) return;
    int64_t i = t / 6;
    int b = (int)(t % 6);
    uint64_t key = inc_keys[i];
    int32_t r = (int32_t)(key >> 32);
    uint32_t e = (uint32_t)key >> 3;
    int32_t col = dofs[(int64_t)e * 6 + b];
    int lo = rowptr[r], hi = rowptr[r + 1];
    int base = lo;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (colidx[mid] < col) lo = mid + 1;
        else hi = mid;
    }
    pos[t] = (uint16_t)(lo - base);
}

// bit 15 of pos[6*i+b]: no earlier incidence of the same row (= lower element, the reference's
// summation order) contributes to the same slot, so the assembly may STORE this contribution
// instead of read-modify-write and needs no zero-fill.  dup is raised when an element lists an
// unknown twice (scalar P1 rides the six-unknown machinery as v0 v1 v2 v0 v1 v2): then the flags are
// not used.
__global__ void k_first_touch(const uint64_t* __restrict__ inc_keys, int64_t n_inc, const int32_t* __restrict__ dofs,
                              const int32_t* __restrict__ incptr, uint16_t* __restrict__ pos, int32_t* __restrict__ dup) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_inc * 6) return;
    const int64_t i = t / 6;
    const int b = (int)(t % 6);
    const uint64_t key = inc_keys[i];
    const int32_t r = (int32_t)(key >> 32);
    const uint32_t e = (uint32_t)key >> 3;
    const int32_t* de = dofs + (int64_t)e * 6;
    const int32_t col = de[b];
    for (int b2 = 0; b2 < 6; ++b2)
        if (b2 != b && de[b2] == col) { *dup = 1; return; }
    bool first = true;
    for (int64_t i2 = incptr[r]; i2 < i && first; ++i2) {
        const int32_t* d2 = dofs + (int64_t)((uint32_t)inc_keys[i2] >> 3) * 6;
        for (int b2 = 0; b2 < 6; ++b2)
            if (d2[b2] == col) first = false;
    }
    if (first) pos[t] |= 0x8000u;
}

// ---- K1 kernels ---------------------------------------------------------------------------
__global__ void k_edge_keys(const int32_t* __restrict__ tris, int64_t n3, uint64_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n3) return;
    int64_t e = i / 3;
    int k = (int)(i % 3);
    uint32_t a = (uint32_t)tris[e * 3 + k], b = (uint32_t)tris[e * 3 + (k + 1) % 3];
    uint32_t lo = min(a, b), hi = max(a, b);
    keys[i] = ((uint64_t)lo << 32) | hi;
    vals[i] = (uint32_t)i;
}

__global__ void k_head_flags(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// runid = inclusive_scan(flags) - 1.  For heads: first[runid] = vals (first appearance, because
// the radix sort is stable and vals ascend inside a run), ids[runid] = runid.
__global__ void k_collect_first(const int32_t* __restrict__ flags, const int32_t* __restrict__ scan_incl,
                                const uint32_t* __restrict__ vals, int64_t n, uint32_t* __restrict__ first,
                                uint32_t* __restrict__ ids) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    int32_t run = scan_incl[i] - 1;
    first[run] = vals[i];
    ids[run] = (uint32_t)run;
}

__global__ void k_rank_from_order(const uint32_t* __restrict__ order, int64_t nu, uint32_t* __restrict__ rank) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < nu) rank[order[j]] = (uint32_t)j;
}

__global__ void k_edges_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                const int32_t* __restrict__ flags, const int32_t* __restrict__ scan_incl,
                                const uint32_t* __restrict__ rank, int64_t n, int32_t* __restrict__ edges,
                                int32_t* __restrict__ edge_idx) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t id = rank[scan_incl[i] - 1];
    edge_idx[vals[i]] = (int32_t)id;
    if (flags[i]) {
        uint64_t k = keys[i];
        edges[2 * (int64_t)id] = (int32_t)(k >> 32);
        edges[2 * (int64_t)id + 1] = (int32_t)(uint32_t)k;
    }
}

// ---- pruning ------------------------------------------------------------------------------
__global__ void k_nz_flags(const double* __restrict__ v, int64_t n, int32_t* __restrict__ f) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) f[i] = (v[i] != 0.0) ? 1 : 0;
}
__global__ void k_prune_scatter(const double* __restrict__ v, const int32_t* __restrict__ col,
                                const int32_t* __restrict__ ex, int64_t n, int32_t* __restrict__ ocol,
                                double* __restrict__ ov) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = v[i];
    if (x != 0.0) {
        int32_t p = ex[i];
        ocol[p] = col[i];
        ov[p] = x;
    }
}
__global__ void k_prune_rowptr(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ ex,
                               int64_t N, int64_t nnz, const int32_t* __restrict__ total,
                               int32_t* __restrict__ optr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > N) return;
    int32_t s = rowptr[r];
    optr[r] = (s < nnz) ? ex[s] : *total;
}
__global__ void k_total(const int32_t* __restrict__ ex, const int32_t* __restrict__ f, int64_t n,
                        int32_t* __restrict__ total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *total = n ? ex[n - 1] + f[n - 1] : 0;
}

static int bits_for(int64_t n) {
    int b = 1;
    while (((int64_t)1 << b) < n) ++b;
    return b;
}

static void sort_keys(uint64_t* in, uint64_t* out, int64_t n, int end_bit, cudaStream_t st) {
    size_t tb = 0;
    WS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, in, out, n, 0, end_bit, st));
    DevBuf<char> tmp(tb, st);
    WS_CUDA(cub::DeviceRadixSort::SortKeys(tmp.get(), tb, in, out, n, 0, end_bit, st));
    count_launch(4);
}

}  // namespace ws

using namespace ws;

extern "C" {

int ws_version(void) { return 100; }
const char* ws_last_error(void) { return ws::g_last_error.c_str(); }
long long ws_launch_count(void) { return ws::g_launches.load(std::memory_order_relaxed); }

int ws_edges_first_appearance(const int32_t* tris, int64_t Ne, int64_t Np, int32_t* edges,
                              int32_t* edge_idx, int64_t* h_Ntt, int device, void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K1 unique edges");
    WS_REQUIRE(Ne >= 0 && Np >= 0 && h_Ntt, "Ne, Np >= 0 and h_Ntt != NULL");
    WS_REQUIRE(Ne * 3 < ((int64_t)1 << 31), "too many elements for 32-bit edge positions");
    *h_Ntt = 0;
    if (Ne == 0) return WS_OK;
    WS_REQUIRE(tris && edges && edge_idx, "null pointer");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = Ne * 3;
    const int T = 256;
    DevBuf<uint64_t> k0(n, st), k1(n, st);
    DevBuf<uint32_t> v0(n, st), v1(n, st);
    k_edge_keys<<<ceil_div(n, T), T, 0, st>>>(tris, n, k0.get(), v0.get());
    int end_bit = 32 + bits_for(Np > 1 ? Np : 2);
    {
        size_t tb = 0;
        WS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.get(), k1.get(), v0.get(), v1.get(), n, 0, end_bit, st));
        DevBuf<char> tmp(tb, st);
        WS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tb, k0.get(), k1.get(), v0.get(), v1.get(), n, 0, end_bit, st));
    }
    DevBuf<int32_t> flags(n, st), scan(n, st);
    k_head_flags<<<ceil_div(n, T), T, 0, st>>>(k1.get(), n, flags.get());
    {
        size_t tb = 0;
        WS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, flags.get(), scan.get(), n, st));
        DevBuf<char> tmp(tb, st);
        WS_CUDA(cub::DeviceScan::InclusiveSum(tmp.get(), tb, flags.get(), scan.get(), n, st));
    }
    int32_t nu32 = 0;
    WS_CUDA(cudaMemcpyAsync(&nu32, scan.get() + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    WS_CUDA(cudaStreamSynchronize(st));
    int64_t nu = nu32;
    // first-appearance position of every unique edge, then order the unique edges by it
    DevBuf<uint32_t> first(nu, st), ids(nu, st), first_s(nu, st), order(nu, st), rank(nu, st);
    k_collect_first<<<ceil_div(n, T), T, 0, st>>>(flags.get(), scan.get(), v1.get(), n, first.get(), ids.get());
    {
        size_t tb = 0;
        int eb = bits_for(n > 1 ? n : 2);
        WS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, first.get(), first_s.get(), ids.get(), order.get(), nu, 0, eb, st));
        DevBuf<char> tmp(tb, st);
        WS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.get(), tb, first.get(), first_s.get(), ids.get(), order.get(), nu, 0, eb, st));
    }
    k_rank_from_order<<<ceil_div(nu, T), T, 0, st>>>(order.get(), nu, rank.get());
    k_edges_scatter<<<ceil_div(n, T), T, 0, st>>>(k1.get(), v1.get(), flags.get(), scan.get(), rank.get(), n, edges, edge_idx);
    WS_CUDA(cudaGetLastError());
    count_launch(5 + 9);
    *h_Ntt = nu;
    WS_API_END
}

int ws_pattern_create(const int32_t* dofs, int64_t Ne, int64_t N, ws_pattern** out, int device,
                      void* stream) {
    WS_API_BEGIN
    WS_NVTX("ws K2 pattern");
    WS_REQUIRE(out, "out == NULL");
    *out = nullptr;
    WS_REQUIRE(Ne >= 0 && N >= 0, "Ne, N >= 0");
    WS_REQUIRE(Ne < ((int64_t)1 << 29), "Ne must be < 2^29 (element id is packed with the local index)");
    WS_REQUIRE(N < ((int64_t)1 << 31), "N must fit int32");
    WS_REQUIRE(Ne * 36 < ((int64_t)1 << 31), "36*Ne must fit int32 (scipy int32 index arrays)");
    DeviceGuard g(device);
    configure_mempool(device);
    cudaStream_t st = (cudaStream_t)stream;
    auto* P = new ws_pattern();
    try {
        P->device = device;
        P->N = N;
        P->Ne = Ne;
        P->rowptr.alloc(N + 1, st);
        P->incptr.alloc(N + 1, st);
        const int T = 256;
        if (Ne == 0) {
            WS_CUDA(cudaMemsetAsync(P->rowptr.get(), 0, (N + 1) * sizeof(int32_t), st));
            WS_CUDA(cudaMemsetAsync(P->incptr.get(), 0, (N + 1) * sizeof(int32_t), st));
            *out = P;
            return WS_OK;
        }
        WS_REQUIRE(dofs, "dofs == NULL");
        int64_t n_inc = Ne * 6, n_pair = Ne * 36;
        int end_bit = 32 + bits_for(N > 1 ? N : 2);
        // --- row incidences, sorted by (row, element, local index)
        DevBuf<uint64_t> ik0(n_inc, st), ik1(n_inc, st);
        k_inc_keys<<<ceil_div(n_inc, T), T, 0, st>>>(dofs, n_inc, ik0.get());
        sort_keys(ik0.get(), ik1.get(), n_inc, end_bit, st);
        P->inc.alloc(n_inc, st);
        k_split_inc<<<ceil_div(n_inc, T), T, 0, st>>>(ik1.get(), n_inc, P->inc.get());
        k_row_bounds<<<ceil_div(N + 1, T), T, 0, st>>>(ik1.get(), n_inc, N, P->incptr.get());
        ik0.release();
        // --- unique (row, col) pairs
        constexpr int NR = ws_pattern::NRANGE;
        DevBuf<int32_t> d_counts(3 + NR, st);
        WS_CUDA(cudaMemsetAsync(d_counts.get(), 0, (3 + NR) * sizeof(int32_t), st));
        int64_t nnz = 0;
        {
            DevBuf<uint64_t> pk0(n_pair, st), pk1(n_pair, st);
            k_pair_keys<<<ceil_div(n_inc, T), T, 0, st>>>(dofs, Ne, pk0.get());
            sort_keys(pk0.get(), pk1.get(), n_pair, end_bit, st);
            size_t tb = 0;
            WS_CUDA(cub::DeviceSelect::Unique(nullptr, tb, pk1.get(), pk0.get(), d_counts.get(), n_pair, st));
            DevBuf<char> tmp(tb, st);
            WS_CUDA(cub::DeviceSelect::Unique(tmp.get(), tb, pk1.get(), pk0.get(), d_counts.get(), n_pair, st));
            int32_t h_nnz = 0;
            WS_CUDA(cudaMemcpyAsync(&h_nnz, d_counts.get(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            WS_CUDA(cudaStreamSynchronize(st));
            nnz = h_nnz;
            P->nnz = nnz;
            P->colidx.alloc(nnz, st);
            k_low32<<<ceil_div(nnz, T), T, 0, st>>>(pk0.get(), nnz, P->colidx.get());
            k_row_bounds<<<ceil_div(N + 1, T), T, 0, st>>>(pk0.get(), nnz, N, P->rowptr.get());
        }
        k_max_rowlen<<<ceil_div(N, T), T, 0, st>>>(P->rowptr.get(), N, d_counts.get() + 1);
        const int nblocks128 = ceil_div(N, 128);
        const int per_group = std::max(1, ceil_div(nblocks128, NR));
        k_max_blockslots<<<ceil_div(nblocks128, T), T, 0, st>>>(P->rowptr.get(), N, 128, per_group, d_counts.get() + 2);
        P->pos.alloc(n_pair, st);
        k_positions<<<ceil_div(n_pair, T), T, 0, st>>>(ik1.get(), n_inc, dofs, P->rowptr.get(), P->colidx.get(), P->pos.get());
        DevBuf<int32_t> d_dup(1, st);
        WS_CUDA(cudaMemsetAsync(d_dup.get(), 0, sizeof(int32_t), st));
        int32_t h_dup = 0;
        int32_t h_mx[2 + NR] = {0};
        WS_CUDA(cudaMemcpyAsync(h_mx, d_counts.get() + 1, (2 + NR) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        WS_CUDA(cudaStreamSynchronize(st));
        WS_CUDA(cudaGetLastError());
        const int32_t h_max = h_mx[0];
        if (h_max < 32768) {     // bit 15 of pos is free
            k_first_touch<<<ceil_div(n_pair, T), T, 0, st>>>(ik1.get(), n_inc, dofs, P->incptr.get(), P->pos.get(), d_dup.get());
            WS_CUDA(cudaMemcpyAsync(&h_dup, d_dup.get(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            WS_CUDA(cudaStreamSynchronize(st));
            P->first_touch = h_dup == 0;
            count_launch(1);
        }
        P->max_row_len = h_max;
        P->max_block_slots = h_mx[1];
        for (int g2 = 0; g2 < NR; ++g2) P->range_block_slots[g2] = h_mx[2 + g2];
        count_launch(9 + 3);
        if (h_max > 65535) throw Error(WS_ERR_ARG, "a matrix row has more than 65535 entries");
    } catch (...) {
        delete P;
        throw;
    }
    *out = P;
    WS_API_END
}

int ws_pattern_query(const ws_pattern* p, int64_t* h_N, int64_t* h_nnz, int64_t* h_Ne) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    if (h_N) *h_N = p->N;
    if (h_nnz) *h_nnz = p->nnz;
    if (h_Ne) *h_Ne = p->Ne;
    WS_API_END
}

int ws_pattern_export(const ws_pattern* p, int32_t* indptr, int32_t* indices, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p && indptr && (indices || p->nnz == 0), "null pointer");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    WS_CUDA(cudaMemcpyAsync(indptr, p->rowptr.get(), (p->N + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    if (p->nnz)
        WS_CUDA(cudaMemcpyAsync(indices, p->colidx.get(), p->nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    WS_API_END
}

int ws_pattern_arrays(const ws_pattern* p, const int32_t** indptr, const int32_t** indices) {
    WS_API_BEGIN
    WS_REQUIRE(p, "pattern == NULL");
    if (indptr) *indptr = p->rowptr.get();
    if (indices) *indices = p->colidx.get();
    WS_API_END
}

int ws_pattern_destroy(ws_pattern* p) {
    WS_API_BEGIN
    if (p) {
        DeviceGuard g(p->device);
        delete p;
    }
    WS_API_END
}

int ws_prune_zeros(const ws_pattern* p, const double* vals, int32_t* out_indptr, int32_t* out_indices,
                   double* out_vals, int64_t* h_nnz_out, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(p && out_indptr && h_nnz_out, "null pointer");
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = p->nnz, N = p->N;
    const int T = 256;
    *h_nnz_out = 0;
    if (n == 0) {
        WS_CUDA(cudaMemsetAsync(out_indptr, 0, (N + 1) * sizeof(int32_t), st));
        return WS_OK;
    }
    WS_REQUIRE(vals && out_indices && out_vals, "null pointer");
    DevBuf<int32_t> f(n, st), ex(n, st), total(1, st);
    k_nz_flags<<<ceil_div(n, T), T, 0, st>>>(vals, n, f.get());
    size_t tb = 0;
    WS_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, f.get(), ex.get(), n, st));
    DevBuf<char> tmp(tb, st);
    WS_CUDA(cub::DeviceScan::ExclusiveSum(tmp.get(), tb, f.get(), ex.get(), n, st));
    k_total<<<1, 32, 0, st>>>(ex.get(), f.get(), n, total.get());
    k_prune_scatter<<<ceil_div(n, T), T, 0, st>>>(vals, p->colidx.get(), ex.get(), n, out_indices, out_vals);
    k_prune_rowptr<<<ceil_div(N + 1, T), T, 0, st>>>(p->rowptr.get(), ex.get(), N, n, total.get(), out_indptr);
    int32_t h = 0;
    WS_CUDA(cudaMemcpyAsync(&h, total.get(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    WS_CUDA(cudaStreamSynchronize(st));
    WS_CUDA(cudaGetLastError());
    count_launch(6);
    *h_nnz_out = h;
    WS_API_END
}

}  // extern "C"
