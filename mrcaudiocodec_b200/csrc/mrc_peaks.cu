// mrc_peaks.cu -- micro-benchmarks for the roofline denominators that MEASURED_PEAKS.json does not carry:
// the FP64 FMA pipe, the FP32 FMA pipe and the MUFU (ex2) pipe of this GPU, under the clocks it actually runs at.
// (The encode hot path is bound by these pipes, not by HBM: SURVEY.md §8d.)
#include "mrc_internal.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_chain(T* out, int iters, T a, T b) {
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void __launch_bounds__(256) ex2_chain(float* out, int iters) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    for (int i = 0; i < iters; ++i) {
        x0 = exp2f(-x0); x1 = exp2f(-x1); x2 = exp2f(-x2); x3 = exp2f(-x3);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (x0 + x1) + (x2 + x3);
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        b[i] = a[i];
}

}  // namespace

// out[0] FP64 TFLOP/s (2 flop per DFMA), out[1] FP32 TFLOP/s, out[2] MUFU.EX2 Gop/s, out[3] copy GB/s (read+write)
int measure_peaks(cudaStream_t st, double* out) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8, threads = 256;
    void* buf = nullptr;
    const size_t copy_bytes = (size_t)1 << 30;
    if (cudaMalloc(&buf, 2 * copy_bytes) != cudaSuccess) return -1;
    cudaMemsetAsync(buf, 1, 2 * copy_bytes, st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    auto best_of = [&](auto launch, int reps) {
        float best = 1e30f;
        for (int r = 0; r < reps; ++r) {
            cudaEventRecord(e0, st);
            launch();
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        return best;
    };
    const int it64 = 4096, it32 = 16384, itx = 8192;
    float t = best_of([&] { fma_chain<double><<<grid, threads, 0, st>>>((double*)buf, it64, 0.999999, 1e-7); }, 4);
    out[0] = 2.0 * 8 * it64 * (double)grid * threads / (t * 1e-3) / 1e12;
    t = best_of([&] { fma_chain<float><<<grid, threads, 0, st>>>((float*)buf, it32, 0.999999f, 1e-7f); }, 4);
    out[1] = 2.0 * 8 * it32 * (double)grid * threads / (t * 1e-3) / 1e12;
    t = best_of([&] { ex2_chain<<<grid, threads, 0, st>>>((float*)buf, itx); }, 4);
    out[2] = 4.0 * itx * (double)grid * threads / (t * 1e-3) / 1e9;
    t = best_of([&] {
        copy_kernel<<<sms * 16, threads, 0, st>>>((const uint4*)buf, (uint4*)((char*)buf + copy_bytes), copy_bytes / 16);
    }, 4);
    out[3] = 2.0 * copy_bytes / (t * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
