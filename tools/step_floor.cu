// step_floor.cu -- the two denominators of the KMC step kernel's roofline, measured on the device:
//   (1) L2 random-sector gather peak: every lane of every resident warp loads random 32-byte
//       entries (one LDG.256, the step kernel's ld_entry) from an L2-resident table of the stencil
//       table's size, many independent loads in flight -> sectors/s and GB/s the L2 delivers to
//       divergent 32-byte gathers;
//   (2) the latency floor of one step's dependency chain: dependent-chain latencies of the
//       operations a step cannot avoid (DFMA, 64-bit shuffle, DMMA m8n8k4, LDS, block barrier of 64
//       threads, REDUX, a warp-wide divergent LDG.256 from L2 and three of them issued together),
//       measured with clock64() on one warp per SM and, for the gathers, under the load of the
//       bench shape (7 warps per SM).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/step_floor tools/step_floor.cu
// Prints one JSON object (profiles/r02_step_floor.json).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void ld256(const double *p, double (&v)[4])
{
#ifdef FLOOR_L1_ALLOCATE   // the step kernel's load before the no-allocate form (profiles/r02_step_floor_l1_allocate.json)
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
#endif
}

__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s; }

// (1) throughput: ILP independent random entries per iteration per lane
template <int ILP>
__global__ void gather_tp_kernel(const double *__restrict__ H, unsigned n_entries, int iters, double *out)
{
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        double v[ILP][4];
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const unsigned idx = (unsigned)(((unsigned long long)lcg(s) * n_entries) >> 32);
            ld256(H + (size_t)idx * 4, v[i]);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += v[i][0] + v[i][3];
    }
    if (acc == 1.2345e300) out[0] = acc;
}

// (2a) dependent divergent gathers: the next index depends on the loaded value (the table holds small
// integers as doubles), NLD loads issued together per link like the tail gathers of a step
template <int NLD>
__global__ void gather_chain_kernel(const double *__restrict__ H, unsigned n_entries, int iters, long long *cycles,
                                    double *out)
{
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    double acc = 0.0;
    unsigned dep = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double v[NLD][4];
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const unsigned idx = (unsigned)(((unsigned long long)(lcg(s) + dep) * n_entries) >> 32);
            ld256(H + (size_t)idx * 4, v[i]);
        }
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < NLD; ++i) sum += v[i][1];
        // the whole warp waits for its slowest lane, as the step does (selection is warp-uniform)
        // the table holds zeros: dep is 0 at run time, but the next addresses depend on the loaded data
        dep = __reduce_add_sync(0xffffffffu, (unsigned)__double2loint(sum));
        acc += sum;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 1.2345e300) out[0] = acc;
}

enum { OP_DFMA, OP_DADD, OP_SHFL64, OP_DMMA, OP_LDS, OP_BAR, OP_REDUX, OP_SHFL32 };

template <int OP>
__global__ void chain_kernel(int iters, long long *cycles, double *out, double seed)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-9, c0 = 0.0, c1 = 0.0;
    unsigned u = threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (OP == OP_DFMA) x = fma(x, 1.0000001, 1e-9);
            if (OP == OP_DADD) x = x + 1e-9;
            if (OP == OP_SHFL64) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9;
            if (OP == OP_SHFL32) u = __shfl_xor_sync(0xffffffffu, u, 1) + 1u;
            if (OP == OP_DMMA) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0), "+d"(c1) : "d"(x), "d"(seed));
            }
            if (OP == OP_LDS) {
                const int i = (int)(__double2loint(x) & 0xff);
                x = sm[i];
            }
            if (OP == OP_BAR) { __syncthreads(); }
            if (OP == OP_REDUX) u = __reduce_add_sync(0xffffffffu, u);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (x + c0 + c1 + u == 1.2345e300) out[0] = x;
}

// DMMA whose A operand depends on the previous result (the scan's stage 1 -> stage 2 dependence)
__global__ void dmma_dep_kernel(int iters, long long *cycles, double *out, double seed)
{
    double a = seed + threadIdx.x * 1e-9;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            double c0 = 0.0, c1 = 0.0;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(seed));
            a = c0;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (a == 1.2345e300) out[0] = a;
}

static double median(std::vector<long long> v)
{
    std::sort(v.begin(), v.end());
    return (double)v[v.size() / 2];
}

int main(int argc, char **argv)
{
    const double table_mb = argc > 1 ? atof(argv[1]) : 31.6;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int n_sm = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const unsigned n_entries = (unsigned)(table_mb * 1e6 / 32);
    double *H, *out;
    long long *cyc;
    CK(cudaMalloc(&H, (size_t)n_entries * 32));
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&cyc, sizeof(long long) * 4096));
    CK(cudaMemset(H, 0, (size_t)n_entries * 32));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_nominal\": %.0f, \"table_mb\": %.1f,\n", prop.name, n_sm,
           clk_khz / 1e3, table_mb);

    // ---- (1) L2 random-sector gather peak ----
    printf(" \"l2_gather_peak\": [\n");
    double best = 0;
    const int warps_list[] = {2, 4, 8, 16, 32, 64};
    bool first = true;
    for (int wi = 0; wi < 6; ++wi) {
        const int warps = warps_list[wi];
        const int bs = 256, ctas = n_sm * warps * 32 / bs > 0 ? n_sm * warps * 32 / bs : n_sm;
        const int threads_per_cta = warps * 32 >= bs ? bs : warps * 32;
        const int grid = warps * 32 >= bs ? ctas : n_sm;
        for (int ilp = 4; ilp <= 16; ilp *= 2) {
            const int iters = 2000 / ilp * 4;
            auto launch = [&] {
                if (ilp == 4) gather_tp_kernel<4><<<grid, threads_per_cta>>>(H, n_entries, iters, out);
                else if (ilp == 8) gather_tp_kernel<8><<<grid, threads_per_cta>>>(H, n_entries, iters, out);
                else gather_tp_kernel<16><<<grid, threads_per_cta>>>(H, n_entries, iters, out);
            };
            launch();
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double sectors = (double)grid * threads_per_cta * iters * ilp;
            const double gbs = sectors * 32 / (ms * 1e-3) / 1e9;
            if (gbs > best) best = gbs;
            printf("%s  {\"warps_per_sm\": %d, \"loads_in_flight_per_lane\": %d, \"ms\": %.3f, \"gsectors_per_s\": %.2f, \"gb_per_s\": %.1f}",
                   first ? "" : ",\n", warps, ilp, ms, sectors / (ms * 1e-3) / 1e9, gbs);
            first = false;
        }
    }
    printf("\n ],\n \"l2_gather_peak_gb_per_s\": %.1f,\n", best);

    // ---- (2) latencies ----
    std::vector<long long> h(4096);
    auto med_cycles = [&](int n) {
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * n, cudaMemcpyDeviceToHost));
        return median(std::vector<long long>(h.begin(), h.begin() + n));
    };
    const int it_c = 2000;
#define CHAIN(OP, NAME, THREADS)                                                                  \
    chain_kernel<OP><<<n_sm, THREADS>>>(it_c, cyc, out, 1.0);                                      \
    CK(cudaDeviceSynchronize());                                                                  \
    printf(" \"lat_%s_cycles\": %.1f,\n", NAME, med_cycles(n_sm) / (it_c * 16.0));
    CHAIN(OP_DFMA, "dfma", 32)
    CHAIN(OP_DADD, "dadd", 32)
    CHAIN(OP_SHFL32, "shfl32_plus_iadd", 32)
    CHAIN(OP_SHFL64, "shfl64_plus_dadd", 32)
    CHAIN(OP_DMMA, "dmma_m8n8k4_accumulate_chain", 32)
    CHAIN(OP_LDS, "lds_dependent", 32)
    CHAIN(OP_BAR, "bar_sync_64_threads", 64)
    CHAIN(OP_BAR, "bar_sync_128_threads", 128)
    CHAIN(OP_REDUX, "redux_add", 32)
    dmma_dep_kernel<<<n_sm, 32>>>(it_c, cyc, out, 1.0);
    CK(cudaDeviceSynchronize());
    printf(" \"lat_dmma_m8n8k4_operand_chain_cycles\": %.1f,\n", med_cycles(n_sm) / (it_c * 16.0));

    // dependent warp-wide divergent gathers: alone (1 warp per SM) and under the bench load (7 warps per SM)
    for (int load = 0; load < 2; ++load) {
        const int grid = load ? n_sm * 7 : n_sm;
        const int it_g = 4000;
        gather_chain_kernel<1><<<grid, 32>>>(H, n_entries, it_g, cyc, out);
        CK(cudaDeviceSynchronize());
        gather_chain_kernel<1><<<grid, 32>>>(H, n_entries, it_g, cyc, out);
        CK(cudaDeviceSynchronize());
        printf(" \"lat_gather_1x32_sectors_%s_cycles\": %.1f,\n", load ? "7_warps_per_sm" : "1_warp_per_sm",
               med_cycles(grid) / (double)it_g);
        gather_chain_kernel<3><<<grid, 32>>>(H, n_entries, it_g, cyc, out);
        CK(cudaDeviceSynchronize());
        gather_chain_kernel<3><<<grid, 32>>>(H, n_entries, it_g, cyc, out);
        CK(cudaDeviceSynchronize());
        printf(" \"lat_gather_3x32_sectors_%s_cycles\": %.1f%s\n", load ? "7_warps_per_sm" : "1_warp_per_sm",
               med_cycles(grid) / (double)it_g, load ? "" : ",");
    }
    printf("}\n");
    return 0;
}
