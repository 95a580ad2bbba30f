// microbench.cu -- FP64 FMA issue-rate probe (roofline denominator for the QP / B&B kernel; the
// driver's MEASURED_PEAKS.json only holds HBM copy and bf16 GEMM).
#include <cuda_runtime.h>

#include "hvp_internal.h"

namespace hvp {

__global__ void __launch_bounds__(256) fp64_fma_kernel(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) sink[threadIdx.x & 1023] = s;   // never true; keeps the chains alive
}

cudaError_t launch_fp64_microbench(int iters, double* sink, int* blocks, int* threads, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *blocks = sms * 8;
    *threads = 256;
    fp64_fma_kernel<<<*blocks, *threads, 0, stream>>>(iters, sink);
    return cudaGetLastError();
}

// shared-memory read bandwidth probe (the denominator for the smem traffic of the QP kernels):
// every thread streams conflict-free 16-byte loads from a 32 KB tile; 4 independent accumulators
__global__ void __launch_bounds__(256) smem_read_kernel(int iters, double* sink) {
    __shared__ double2 tile[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) tile[i] = make_double2(i * 1e-9, 1.0);
    __syncthreads();
    double2 a0 = make_double2(0, 0), a1 = a0, a2 = a0, a3 = a0;
    int j = threadIdx.x;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        const double2 v0 = tile[j], v1 = tile[(j + 256) & 2047], v2 = tile[(j + 512) & 2047], v3 = tile[(j + 768) & 2047];
        a0.x += v0.x; a0.y += v0.y; a1.x += v1.x; a1.y += v1.y; a2.x += v2.x; a2.y += v2.y; a3.x += v3.x; a3.y += v3.y;
        j = (j + 1056) & 2047;      // period 64 in i: the addresses keep changing, nothing is loop invariant
    }
    const double s = (a0.x + a0.y) + (a1.x + a1.y) + (a2.x + a2.y) + (a3.x + a3.y);
    if (s == 123.456) sink[threadIdx.x & 1023] = s;
}

cudaError_t launch_smem_microbench(int iters, double* sink, int* blocks, int* threads, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *blocks = sms * 4;
    *threads = 256;
    smem_read_kernel<<<*blocks, *threads, 0, stream>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace hvp
