// Shared helpers for the wavesolve_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / Compute, no link dependency

#include "../../include/wavesolve_b200.h"

namespace ws {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& s);

#define WS_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            char buf__[512];                                                                 \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,      \
                     cudaGetErrorString(e__));                                               \
            throw ws::Error(WS_ERR_CUDA, buf__);                                             \
        }                                                                                    \
    } while (0)

#define WS_REQUIRE(cond, msg)                                                                \
    do {                                                                                     \
        if (!(cond)) throw ws::Error(WS_ERR_ARG, std::string("argument error: ") + (msg));   \
    } while (0)

#define WS_API_BEGIN try {
#define WS_API_END                                                                           \
    }                                                                                        \
    catch (const ws::Error& e) {                                                             \
        ws::set_last_error(e.what());                                                        \
        return e.code;                                                                       \
    }                                                                                        \
    catch (const std::exception& e) {                                                        \
        ws::set_last_error(e.what());                                                        \
        return WS_ERR_CUDA;                                                                  \
    }                                                                                        \
    catch (...) {                                                                            \
        ws::set_last_error("unknown exception");                                             \
        return WS_ERR_CUDA;                                                                  \
    }                                                                                        \
    return WS_OK;

// Stream-ordered device buffer (cudaMallocAsync: no implicit device synchronisation).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t st = nullptr;
    DevBuf() = default;
    DevBuf(size_t n_, cudaStream_t s) { alloc(n_, s); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p; n = o.n; st = o.st;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
    void alloc(size_t n_, cudaStream_t s) {
        release();
        n = n_;
        st = s;
        if (n) WS_CUDA(cudaMallocAsync((void**)&p, n * sizeof(T), s));
    }
    void release() {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    T* get() const { return p; }
    size_t bytes() const { return n * sizeof(T); }
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        WS_CUDA(cudaGetDevice(&prev));
        if (dev != prev) WS_CUDA(cudaSetDevice(dev));
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// cudaFuncSetAttribute is a per-device setting: run a (idempotent) configuration once per device
// the calling thread is bound to.  Two threads racing on the first use both run it -- harmless.
struct PerDeviceOnce {
    std::atomic<uint64_t> done{0};
    template <class F>
    void operator()(F&& f) {
        int dev = 0;
        WS_CUDA(cudaGetDevice(&dev));
        const uint64_t bit = 1ull << (dev & 63);
        if (done.load(std::memory_order_acquire) & bit) return;
        f();
        done.fetch_or(bit, std::memory_order_release);
    }
};

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// NVTX range around a phase of the hot path (pattern / assemble / symbolic / factor / iterate)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define WS_NVTX(name) ws::NvtxRange nvtx_range__(name)

// keep freed stream-ordered allocations cached in the device's default pool (the solver
// allocates ~12 GB per 1M-element problem; returning it to the OS after every solve costs
// ~100 ms per call).  Idempotent, called by the handle constructors.
void configure_mempool(int device);

// global launch counter (bench.py reports how many of OUR kernels ran in the timed region)
void count_launch(int n = 1);

#ifdef __CUDACC__
// fp64 tensor-core MMA (DMMA m8n8k4; tcgen05 has no fp64 kind, so this is the sm_100a fp64 tensor path).
// Fragment layout (g = lane >> 2, tq = lane & 3): a = A[g][tq], b = B[tq][g], c0/c1 = C[g][2 tq], C[g][2 tq + 1].
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
#endif

}  // namespace ws

// ---------------------------------------------------------------------------------------
// pattern object shared by the assembly / SpMV / solver translation units
// ---------------------------------------------------------------------------------------
struct ws_pattern {
    int device = 0;
    int64_t N = 0;     // rows
    int64_t Ne = 0;    // elements (as assembled, material-major, duplicates allowed)
    int64_t nnz = 0;
    int max_row_len = 0;
    int max_block_slots = 0;   // largest number of entries in a block of 128 consecutive rows
    static constexpr int NRANGE = 16;
    int range_block_slots[NRANGE] = {0};   // the same maximum per sixteenth of the 128-row blocks
    ws::DevBuf<int32_t> rowptr;   // [N+1]
    ws::DevBuf<int32_t> colidx;   // [nnz]
    // row-incidence lists: for row r, entries incptr[r]..incptr[r+1] of inc[] hold
    // (element << 3 | local index a), sorted by element (= reference summation order);
    // pos[6*i + b] is the position inside row r of column dofs[e*6+b].
    ws::DevBuf<int32_t> incptr;   // [N+1]
    ws::DevBuf<uint32_t> inc;     // [6*Ne]
    ws::DevBuf<uint16_t> pos;     // [36*Ne]  bit 15: this is the first contribution to its slot
    bool first_touch = false;     // the bit-15 flags are valid (no element lists an unknown twice)
    // cache of ws_vector_T (vector problem): edge lengths and the node -> incident-edge lists of the
    // discrete gradient, built on the first application for a given (edge table, coordinates)
    mutable const void* vt_edges = nullptr;
    mutable const void* vt_xy = nullptr;
    mutable int64_t vt_Ntt = -1;
    mutable ws::DevBuf<double> vt_len;        // [Ntt]
    mutable ws::DevBuf<int32_t> vt_nptr;      // [Nzz + 1]
    mutable ws::DevBuf<int32_t> vt_nadj;      // [2*Ntt]  edge id << 1 | (node is the edge's head)
};
