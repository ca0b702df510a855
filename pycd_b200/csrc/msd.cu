// msd.cu -- time-origin-averaged squared displacement on sm_100a.
// Replaces the lag loop of Analysis.compute_msd (PyCD/core.py:2996-3022):
//   sd[traj,tau,c] = mean_{t=0..n_path-tau-1} |r(t+tau,c) - r(t,c)|^2   (tau >= 1; tau = 0 -> 0)
//   species_avg[traj,tau,type] = mean over the carriers of the type
// One CTA per (trajectory, carrier).  Lags are processed in windows of MSD_LAGW = 2048 (one lag per
// thread and pass), time origins in tiles of MSD_TILE; per (window, tile) the carrier's positions are
// staged through shared memory as two segments -- the origins [t0, t0 + TILE) and the targets
// [t0 + lag0, t0 + lag0 + TILE + LAGW) -- so the footprint does not depend on n_msd (any n_msd <=
// n_path works, like the reference) and every position is read from HBM/L2 twice per lag window instead
// of once per lag.
#include "common.cuh"

#include <algorithm>

namespace pycd {

constexpr int MSD_THREADS = 256;
constexpr int MSD_TILE = 2048;
constexpr int MSD_MAX_LAGS = 8;                          // lags per thread and window
constexpr int MSD_LAGW = MSD_THREADS * MSD_MAX_LAGS;     // lags per window

__global__ void __launch_bounds__(MSD_THREADS)
msd_sd_kernel(const double *__restrict__ unwrapped, long long n_path, int C, long long n_msd, int lag_w,
              double scale, double *__restrict__ sd /* [n_traj][n_msd][C] */)
{
    extern __shared__ __align__(16) double s_pos[];  // origins [3][MSD_TILE], targets [3][MSD_TILE + lag_w]
    const long long traj = blockIdx.x / C;
    const int c = blockIdx.x % C;
    const int tid = threadIdx.x;
    const int span_b = MSD_TILE + lag_w;
    double *ax = s_pos, *ay = ax + MSD_TILE, *az = ay + MSD_TILE;
    double *bx = az + MSD_TILE, *by = bx + span_b, *bz = by + span_b;
    const double *base = unwrapped + traj * n_path * 3 * C + 3 * c;

    // per-thread lags of a window: tau = lag0 + tid + i*MSD_THREADS
    for (long long lag0 = 1; lag0 < n_msd; lag0 += MSD_LAGW) {
        double acc[MSD_MAX_LAGS];
#pragma unroll
        for (int i = 0; i < MSD_MAX_LAGS; ++i) acc[i] = 0.0;
        for (long long t0 = 0; t0 + lag0 < n_path; t0 += MSD_TILE) {
            const long long n_orig = min((long long)MSD_TILE, n_path - t0);
            const long long n_tgt = min((long long)span_b, n_path - t0 - lag0);
            __syncthreads();
            for (long long r = tid; r < n_orig; r += MSD_THREADS) {
                const double *p = base + (t0 + r) * 3 * C;
                ax[r] = p[0] * scale;  // position_array * dist_conversion, core.py:2990-2995
                ay[r] = p[1] * scale;
                az[r] = p[2] * scale;
            }
            for (long long r = tid; r < n_tgt; r += MSD_THREADS) {
                const double *p = base + (t0 + lag0 + r) * 3 * C;
                bx[r] = p[0] * scale;
                by[r] = p[1] * scale;
                bz[r] = p[2] * scale;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < MSD_MAX_LAGS; ++i) {
                const int j = tid + i * MSD_THREADS;   // tau - lag0
                const long long tau = lag0 + j;
                if (j >= lag_w || tau >= n_msd) break;
                // origins t in [t0, t0+n_orig) with t + tau < n_path
                const long long lim = min(n_orig, n_path - tau - t0);
                double s = 0.0;
                for (long long r = 0; r < lim; ++r) {
                    const double dx = bx[r + j] - ax[r], dy = by[r + j] - ay[r], dz = bz[r + j] - az[r];
                    s += dx * dx + dy * dy + dz * dz;
                }
                acc[i] += s;
            }
        }
#pragma unroll
        for (int i = 0; i < MSD_MAX_LAGS; ++i) {
            const long long tau = lag0 + tid + (long long)i * MSD_THREADS;
            if (tid + i * MSD_THREADS < lag_w && tau < n_msd)
                sd[(traj * n_msd + tau) * C + c] = acc[i] / (double)(n_path - tau);
        }
    }
    if (tid == 0) sd[(traj * n_msd) * C + c] = 0.0;
}

// species_avg[traj,tau,type] = mean_c sd[traj,tau,c] over the type's carriers (fixed order)
__global__ void msd_species_kernel(const double *__restrict__ sd, long long n_rows /* n_traj*n_msd */,
                                   int C, const int *__restrict__ type_offsets, int n_types,
                                   double *__restrict__ out)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= n_rows * n_types) return;
    const long long r = idx / n_types;
    const int ty = (int)(idx - r * n_types);
    const int c0 = type_offsets[ty], c1 = type_offsets[ty + 1];
    double s = 0.0;
    for (int c = c0; c < c1; ++c) s += sd[r * C + c];
    out[idx] = s / (double)(c1 - c0);
}

}  // namespace pycd

using namespace pycd;

extern "C" int pycd_msd(pycd_ctx *ctx, const double *unwrapped, int64_t n_traj, int64_t n_path,
                        int32_t n_carriers, int64_t n_msd, double scale, const int32_t *type_offsets,
                        int32_t n_types, double *sd_species, double *sd_carrier) {
    return guarded([&] {
        NvtxRange nvtx("pycd.msd");
        PYCD_REQUIRE(ctx && unwrapped && type_offsets && sd_species, "NULL argument");
        PYCD_REQUIRE(n_traj > 0 && n_path > 0 && n_carriers > 0 && n_types > 0, "bad sizes");
        PYCD_REQUIRE(n_msd >= 1 && n_msd <= n_path, "need 1 <= n_msd <= n_path (msd_t_final <= t_final)");
        PYCD_REQUIRE(n_traj * n_carriers < (1ll << 31), "too many (trajectory, carrier) pairs");
        DeviceGuard g(ctx);
        cudaStream_t s = ctx->stream;
        const int C = n_carriers;
        std::vector<int> toff(n_types + 1);
        PYCD_CUDA(cudaMemcpy(toff.data(), type_offsets, sizeof(int) * (n_types + 1), cudaMemcpyDefault));
        PYCD_REQUIRE(toff[0] == 0 && toff[n_types] == C, "type_offsets must span [0, C]");
        for (int k = 0; k < n_types; ++k) PYCD_REQUIRE(toff[k + 1] > toff[k], "empty carrier type");
        InBuf<double> pos;
        pos.bind(unwrapped, (size_t)n_traj * n_path * 3 * C, s);
        DevBuf<int> d_toff;
        d_toff.alloc(n_types + 1);
        PYCD_CUDA(cudaMemcpyAsync(d_toff.p, toff.data(), sizeof(int) * (n_types + 1), cudaMemcpyHostToDevice, s));
        OutBuf<double> o_sp, o_c;
        o_sp.bind(sd_species, (size_t)n_traj * n_msd * n_types);
        DevBuf<double> sd_tmp;
        double *sd_dev;
        if (sd_carrier) {
            o_c.bind(sd_carrier, (size_t)n_traj * n_msd * C);
            sd_dev = o_c.dev();
        } else {
            sd_tmp.alloc((size_t)n_traj * n_msd * C);
            sd_dev = sd_tmp.p;
        }
        // lag window: all lags at once when they fit one window (the shipped examples: n_msd = 501)
        const int lag_w = (int)std::min<long long>(MSD_LAGW, std::max<long long>(n_msd - 1, 1));
        const size_t smem = sizeof(double) * 3 * (size_t)(2 * MSD_TILE + lag_w);   // <= 147 KB
        PYCD_CUDA(cudaFuncSetAttribute(msd_sd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelTimer t(ctx, KC_MSD);
        msd_sd_kernel<<<(unsigned)(n_traj * C), MSD_THREADS, smem, s>>>(pos.p, n_path, C, n_msd, lag_w, scale,
                                                                       sd_dev);
        check_launch(ctx, "msd_sd_kernel");
        const long long n_out = n_traj * n_msd * n_types;
        msd_species_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, s>>>(sd_dev, n_traj * n_msd, C, d_toff.p,
                                                                           n_types, o_sp.dev());
        check_launch(ctx, "msd_species_kernel");
        t.stop(2);
        o_sp.finish(s);
        o_c.finish(s);
        PYCD_CUDA(cudaStreamSynchronize(s));
        t.read();
    });
}
