// Measured device peaks for the roofline denominators of the fused likelihood
// kernel: MUFU.EX2 issue rate and FP32 FMA rate, measured on the
// same box and clocks as the benchmark.  Pure register-resident loops.
#include "nf_internal.cuh"

namespace {

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int ILP>
__global__ void __launch_bounds__(256) mufu_kernel(float *out, int iters)
{
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = -0.001f * (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = ex2_approx(v[i]) - 1.5f;   // 1 MUFU + 1 FADD
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters)
{
    float v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = 0.001f * (float)(threadIdx.x + i);
    const float a = 0.999f, b = 0.0001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = fmaf(v[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    if (s == 12345.678f) out[0] = s;
}

}  // namespace

extern "C" int nf_measure_peaks(int device, double *mufu_gops, double *ffma_gflops)
{
    if (!mufu_gops || !ffma_gflops) return NF_EINVAL;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return NF_ENODEV;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return (int)e;
    const int grid = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    constexpr int ILP = 8;
    float *out = nullptr;
    if ((e = cudaMalloc(&out, 4)) != cudaSuccess) return (int)e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best_m = 0.0, best_f = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        float ms = 0.f;
        cudaEventRecord(e0);
        mufu_kernel<ILP><<<grid, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)grid * threads * iters * ILP;
        if (ms > 0 && ops / ms * 1e-6 > best_m) best_m = ops / ms * 1e-6;
        cudaEventRecord(e0);
        ffma_kernel<ILP><<<grid, threads>>>(out, iters * 4);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * (double)grid * threads * iters * 4 * ILP;
        if (ms > 0 && fl / ms * 1e-6 > best_f) best_f = fl / ms * 1e-6;
    }
    e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (prev >= 0) cudaSetDevice(prev);
    *mufu_gops = best_m;
    *ffma_gflops = best_f;
    return (int)e;
}
