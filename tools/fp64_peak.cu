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


// ---- what the Ewald kernel actually issues: 4x4 accumulator tiles with distinct fragments, fragments
// from shared memory, and scalar FP64 work (the phase recurrence) sharing the pipe ----
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: 4x4 tile, fragments in registers.  MODE 1: fragments re-read from shared memory every k4 slice
// (8 LDS.64 per 16 DMMA, the kernel's ratio).  NF: dependent scalar-FP64 complex-multiply steps
// (2 DMUL + 2 DFMA each) issued by the same warp per 16 DMMAs.
template <int MODE, int NF>
__global__ void dmma_tile_kernel(double *out, int iters, double seed)
{
    __shared__ double frag[2][8][32];
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < 8; ++i) { frag[0][i][lane] = seed + i + lane * 1e-3; frag[1][i][lane] = seed * 0.5 + i; }
    __syncthreads();
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = frag[0][i][lane]; b[i] = frag[0][4 + i][lane]; }
    double c = 1.0, s = 0.0;
    const double c3 = 0.9999995, s3 = 0.001;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) {
            const volatile double *f = &frag[it & 1][0][0];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = f[i * 32 + lane]; b[i] = f[(4 + i) * 32 + lane]; }
        }
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            const double cn = c * c3 - s * s3;
            s = fma(s, c3, c * s3);
            c = cn;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    double t = c + s;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) t += acc[i][j][0] + acc[i][j][1];
    if (t == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

// warp-specialised mix: warps [0, NGEN) run NCH independent dependent complex-multiply chains,
// the other warps the 4x4 DMMA tile; both loop `iters` times (the DMMA warps do 16 DMMA per
// iteration, the scalar warps NSTEP chain steps per iteration).
template <int NGEN, int NSTEP>
__global__ void dmma_split_kernel(double *out, int iters, double seed)
{
    const int wid = threadIdx.x >> 5;
    if (wid < NGEN) {
        double c = 1.0 + threadIdx.x * 1e-9, s = 0.0;
        const double c3 = 0.9999995, s3 = 0.001;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < NSTEP; ++q) {
                const double cn = c * c3 - s * s3;
                s = fma(s, c3, c * s3);
                c = cn;
            }
        }
        if (c + s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = c;
        return;
    }
    double acc[4][4][2], a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = seed + i + threadIdx.x * 1e-9; b[i] = seed * 0.5 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    double t = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) t += acc[i][j][0] + acc[i][j][1];
    if (t == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = t;
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

    // DMMA flops only (the scalar work is overhead, as in the Ewald kernel)
    for (int bps = 1; bps <= 2; ++bps) {
        const int t = 256, grid = sms * bps, it2 = iters / 4;
        auto rep2 = [&](const char *name, double warps, double ms) {
            printf(",\n{\"kernel\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"dmma_tflops\": %.2f}", name, t, bps,
                   2.0 * 256 * 16 * it2 * warps * grid / (ms * 1e-3) / 1e12);
        };
        double ms;
        ms = time_ms([&] { dmma_tile_kernel<0, 0><<<grid, t>>>(out, it2, 1.5); }); rep2("tile4x4_regs", 8, ms);
        ms = time_ms([&] { dmma_tile_kernel<1, 0><<<grid, t>>>(out, it2, 1.5); }); rep2("tile4x4_lds", 8, ms);
        ms = time_ms([&] { dmma_tile_kernel<0, 1><<<grid, t>>>(out, it2, 1.5); }); rep2("tile4x4_regs+1cmul/16", 8, ms);
        ms = time_ms([&] { dmma_tile_kernel<0, 2><<<grid, t>>>(out, it2, 1.5); }); rep2("tile4x4_regs+2cmul/16", 8, ms);
        ms = time_ms([&] { dmma_tile_kernel<0, 4><<<grid, t>>>(out, it2, 1.5); }); rep2("tile4x4_regs+4cmul/16", 8, ms);
        ms = time_ms([&] { dmma_tile_kernel<1, 2><<<grid, t>>>(out, it2, 1.5); }); rep2("tile4x4_lds+2cmul/16", 8, ms);
        ms = time_ms([&] { dmma_split_kernel<1, 4><<<grid, t>>>(out, it2, 1.5); }); rep2("split_7dmma+1gen(4 steps)", 7, ms);
        ms = time_ms([&] { dmma_split_kernel<1, 16><<<grid, t>>>(out, it2, 1.5); }); rep2("split_7dmma+1gen(16 steps)", 7, ms);
        ms = time_ms([&] { dmma_split_kernel<2, 8><<<grid, t>>>(out, it2, 1.5); }); rep2("split_6dmma+2gen(8 steps)", 6, ms);
        ms = time_ms([&] { dmma_split_kernel<4, 4><<<grid, t>>>(out, it2, 1.5); }); rep2("split_4dmma+4gen(4 steps)", 4, ms);
    }
    printf("\n]}\n");
    return 0;
}
