// Measures what the FP64 pipe of this GPU sustains (warp-level DFMA issue rate) so that the
// fused flux kernel can be placed against an FP64 roofline as well as the HBM one.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int n = 0; n < iters; ++n) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
double run(int ctas_per_sm, int threads, int iters, int nsm, int mhz, double* d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = nsm * ctas_per_sm;
    dfma_kernel<CHAINS><<<grid, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dfma_kernel<CHAINS><<<grid, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)grid * threads * CHAINS * iters;
    const double per_clk_sm = fmas / (ms * 1e-3) / nsm / (mhz * 1e6);
    printf("chains=%d ctas/sm=%d threads=%d warps/sm=%d : %.3f ms  %.2f TFLOP/s  %.1f DFMA lanes/clk/SM (at %d MHz)\n", CHAINS,
           ctas_per_sm, threads, ctas_per_sm * threads / 32, ms, 2 * fmas / (ms * 1e-3) / 1e12, per_clk_sm, mhz);
    return per_clk_sm;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int mhz = p.clockRate / 1000;
    printf("%s: %d SMs, %d MHz\n", p.name, p.multiProcessorCount, mhz);
    double* d;
    cudaMalloc(&d, 8);
    const int nsm = p.multiProcessorCount;
    run<1>(1, 32, 200000, nsm, mhz, d);      // one warp, one chain: DFMA latency = 32 / lanes
    run<1>(1, 128, 200000, nsm, mhz, d);     // one warp per scheduler
    run<2>(1, 128, 200000, nsm, mhz, d);
    run<4>(1, 128, 100000, nsm, mhz, d);
    run<8>(1, 128, 100000, nsm, mhz, d);
    run<8>(2, 256, 50000, nsm, mhz, d);      // 16 warps/SM like the flux kernel
    run<8>(4, 256, 50000, nsm, mhz, d);
    run<4>(2, 256, 50000, nsm, mhz, d);
    run<2>(2, 256, 50000, nsm, mhz, d);
    run<1>(2, 256, 50000, nsm, mhz, d);
    return 0;
}
