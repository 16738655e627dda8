// K12, host side of the gather: device -> pageable host memory for GB-sized payloads.
//
// The gathered eigenvectors of a sweep (2 GB for 256 problems of 100k unknowns) land on ONE host.  Page-locking a
// landing buffer of that size costs ~1 s per GB (more than the transfer), a pageable cudaMemcpy is bound by the
// driver's single staging thread, and a Python-level double-buffered copy by numpy's single-threaded memcpy into
// freshly mapped pages (3.3 GB/s measured).  Here the payload goes through two reusable page-locked staging blocks:
// the DMA of chunk i+1 runs while a few host threads copy chunk i out (the page faults of a fresh destination
// parallelise across threads).  Replaces the pickle / np.save of results in lantern-demo.ipynb cell 4 as the place
// where sweep results reach host memory.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "ws_common.cuh"

namespace {

constexpr size_t STAGE_BYTES = (size_t)64 << 20;
std::mutex g_stage_mutex;
void* g_stage[2] = {nullptr, nullptr};      // kept for the life of the process (128 MB of pinned memory)

void copy_out(char* dst, const char* src, size_t n, int nthreads) {
    if (nthreads <= 1 || n < ((size_t)4 << 20)) {
        std::memcpy(dst, src, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((n + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    for (int t = 0; t < nthreads; ++t) {
        const size_t a = std::min(n, (size_t)t * per), b = std::min(n, a + per);
        if (b > a) th.emplace_back([=] { std::memcpy(dst + a, src + a, b - a); });
    }
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" int ws_copy_to_host(const void* d_src, void* h_dst, int64_t nbytes, int nthreads, int device, void* stream) {
    WS_API_BEGIN
    WS_REQUIRE(nbytes >= 0 && (nbytes == 0 || (d_src && h_dst)), "bad arguments");
    if (nbytes == 0) return WS_OK;
    ws::DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (nthreads <= 0) nthreads = (int)std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency() / 4));
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    for (int i = 0; i < 2; ++i)
        if (!g_stage[i]) WS_CUDA(cudaMallocHost(&g_stage[i], STAGE_BYTES));
    cudaEvent_t ev[2];
    WS_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    WS_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    const char* src = (const char*)d_src;
    char* dst = (char*)h_dst;
    const size_t n = (size_t)nbytes;
    size_t pend_off = 0, pend_n = 0;
    int pend_buf = -1, j = 0;
    try {
        for (size_t off = 0; off < n; off += STAGE_BYTES, ++j) {
            const size_t cn = std::min(STAGE_BYTES, n - off);
            const int b = j & 1;
            WS_CUDA(cudaMemcpyAsync(g_stage[b], src + off, cn, cudaMemcpyDeviceToHost, st));
            WS_CUDA(cudaEventRecord(ev[b], st));
            if (pend_buf >= 0) {
                WS_CUDA(cudaEventSynchronize(ev[pend_buf]));
                copy_out(dst + pend_off, (const char*)g_stage[pend_buf], pend_n, nthreads);
            }
            pend_off = off; pend_n = cn; pend_buf = b;
        }
        WS_CUDA(cudaEventSynchronize(ev[pend_buf]));
        copy_out(dst + pend_off, (const char*)g_stage[pend_buf], pend_n, nthreads);
    } catch (...) {
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
        throw;
    }
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    WS_API_END
}
