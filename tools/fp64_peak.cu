// fp64_peak.cu -- measures the FP64 DFMA / DMMA issue ceiling of the device: register-only
// loops (no memory), swept over resident warps per SM and independent accumulators per thread.
// The best figure is the measured denominator for the Ewald kernel's roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double b, double c)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// a*b+acc with distinct a, b per accumulator row/col like a register-tiled GEMM (TM x TN)
template <int TM, int TN>
__global__ void dfma_tile_kernel(double *out, int iters, double seed)
{
    double acc[TM][TN], a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = seed + i + threadIdx.x * 1e-9;
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = seed * 0.5 + j;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] += 1e-12;   // keep the loop from being hoisted
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) s += acc[i][j];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma_kernel(double *out, int iters, double seed)
{
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = 0; c1[i] = 0; }
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, sizeof(double) * 1024 * 1024 * 16);
    const int iters = 20000;
    printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
    const int threads[] = {128, 256, 512, 1024};
    bool first = true;
    for (int t : threads) {
        for (int bps = 1; bps <= 2; ++bps) {
            if (t * bps > 2048) continue;
            const int grid = sms * bps;
            auto rep = [&](const char *name, double flops, double ms) {
                printf("%s{\"kernel\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"tflops\": %.2f}", first ? "" : ",\n",
                       name, t, bps, flops / (ms * 1e-3) / 1e12);
                first = false;
            };
            double ms = time_ms([&] { dfma_kernel<8><<<grid, t>>>(out, iters, 1.0000001, 1e-9); });
            rep("dfma_ilp8", 2.0 * 8 * iters * (double)t * grid, ms);
            ms = time_ms([&] { dfma_kernel<16><<<grid, t>>>(out, iters, 1.0000001, 1e-9); });
            rep("dfma_ilp16", 2.0 * 16 * iters * (double)t * grid, ms);
            if (t <= 512) {
                ms = time_ms([&] { dfma_tile_kernel<8, 4><<<grid, t>>>(out, iters / 4, 1.5); });
                rep("dfma_tile8x4", 2.0 * 32 * (iters / 4) * (double)t * grid, ms);
            }
            if (t <= 256 && bps == 1) {
                ms = time_ms([&] { dfma_tile_kernel<8, 8><<<grid, t>>>(out, iters / 8, 1.5); });
                rep("dfma_tile8x8", 2.0 * 64 * (iters / 8) * (double)t * grid, ms);
            }
            ms = time_ms([&] { dmma_kernel<8><<<grid, t>>>(out, iters / 4, 1.5); });
            rep("dmma_m8n8k4_x8", 2.0 * 256 * 8 * (iters / 4) * (double)(t / 32) * grid, ms);
            if (t <= 512) {
                ms = time_ms([&] { dmma_kernel<16><<<grid, t>>>(out, iters / 4, 1.5); });
                rep("dmma_m8n8k4_x16", 2.0 * 256 * 16 * (iters / 4) * (double)(t / 32) * grid, ms);
            }
        }
    }
    printf("\n]}\n");
    return 0;
}
