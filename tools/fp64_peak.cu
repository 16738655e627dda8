// Measured fp64 peaks of the B200 in the box (the denominators behind DESIGN.md's statements about
// the FP64 pipe): dependent-chain-free DFMA throughput and mma.sync.m8n8k4.f64 (DMMA) throughput.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_peak.cu -o /tmp/fp64_peak && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int kind = 0; kind < 2; ++kind) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (kind == 0) k_dfma<<<blocks, threads>>>(out, iters);
            else k_dmma<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        // DFMA: 2 flop per thread-op; DMMA m8n8k4: 8*8*4*2 = 512 flop per warp-op
        const double ops = (double)blocks * threads * iters * 8;
        const double flops = kind == 0 ? ops * 2 : ops / 32 * 512;
        printf("%s: %.3f ms, %.2f TFLOP/s (%d SMs, %d MHz)\n", kind == 0 ? "DFMA" : "DMMA m8n8k4", best, flops / best / 1e9,
               p.multiProcessorCount, p.clockRate / 1000);
    }
    return 0;
}
