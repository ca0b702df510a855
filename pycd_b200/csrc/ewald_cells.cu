// ewald_cells.cu -- reciprocal-space part of the UNIT-CELL ROWS of the Ewald array under full periodic
// boundaries, factorised by residue classes of the k indices modulo the supercell size.
//
// The reference sums, for every site pair, cos(k.d_ij) over all half-space k vectors of the SUPERCELL
// (PyCD/core.py:853-878): 4 N^2 K_eff flop in the structure-factor form of ewald.cu, 4 n_b N K_eff for the
// rows of unit cell 0.  With sites r = tau_b + R (R = x a1 + y a2 + z a3, a_i = unit-cell vectors) and
// k = n.B (B reciprocal to the supercell s_i a_i):   k.R = 2 pi sum_i n_i R_i / s_i,   which depends on n only
// through the residues p_i = n_i mod s_i.  Hence, for a site b of cell 0 and a site b' of cell R,
//
//   F(b, b', R) = 2 sum_k w_k cos(k.(tau_b' - tau_b) + phi_p(R)),     phi_p(R) = 2 pi sum_i p_i R_i / s_i
//               = 2 sum_p [ cos phi_p(R) C_p(b,b') - sin phi_p(R) S_p(b,b') ]
//   C_p(b,b') = sum_{k in class p} w_k cos(k.(tau_b' - tau_b)),   S_p likewise with sin
//
// i.e. ONE pass over the k vectors for the n_b^2 basis pairs (class sums, kernel A: 4 n_b^2 K_eff flop --
// n_cells times fewer than the rows of unit cell 0, 1000x at Hematite 10x10x10) followed by a discrete Fourier
// transform over the n_cells classes per basis pair (kernel B).  Same sum, different order: agrees with the
// DMMA kernels of ewald.cu to ~1e-15 max|P| (tests/test_gpu_configs.py).  The k vectors are enumerated on the
// device, class by class (no host k list); real-space and self terms come from ewald_finish_kernel as before.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace pycd {

struct FinishParams;   // ewald.cu
void ewald_finish_rows(pycd_ctx *ctx, const pycd_ewald_desc *desc, const double *coords_dev, long long n,
                       long long n_rows, const double *partials, double *out_dev);   // ewald.cu

struct CellParams {
    int nb, sx, sy, sz;
    int kmax[3];
    double B[9];          // supercell reciprocal rows
    double coeff;         // 2 pi / V
    double alpha4, kc2;
};

constexpr int EC_THREADS = 256;
constexpr int EC_KCH = 8;        // k vectors per accumulation sub-chunk

// ---- kernel A: class sums.  grid = (n_classes, pair tiles).  A thread owns a 2 x 4 tile of basis pairs
// (b, b+1) x (b'..b'+3): per k vector 7 shared-memory loads (weight, 2 x 16 bytes of cos/sin of its b pair,
// 4 x 16 bytes of its four b') feed 32 FMAs -- the first version (1 x 4 tile, 11 eight-byte loads per 16 FMAs)
// ran at 86 % of the shared-memory wavefront peak with the FP64 pipe 29 % busy.  When the pairs need fewer
// than 256 threads, `ks` groups of `ipad` threads take the k vectors of a sub-chunk in turn and their partial
// sums are added in group order at the end (fixed order: results do not depend on scheduling).
__global__ void __launch_bounds__(EC_THREADS, 2)
ewald_class_sums_kernel(CellParams P, const double *__restrict__ theta /* [nb][3] = B.tau_b */,
                        double *__restrict__ Cs, double *__restrict__ Ss /* [class][nb][nb] */,
                        unsigned long long *__restrict__ k_count, int ipad, int ks)
{
    extern __shared__ __align__(16) double sm[];
    const int nb = P.nb, nbp = (nb + 3) & ~3;
    double *s_c = sm;                       // [KCH][nbp]  cos(k.tau_b), rows padded with zeros
    double *s_s = s_c + EC_KCH * nbp;       // [KCH][nbp]  sin
    double *s_w = s_s + EC_KCH * nbp;       // [EC_THREADS] weights of the valid k of a batch
    int *s_n = reinterpret_cast<int *>(s_w + EC_THREADS);   // [EC_THREADS][3]
    __shared__ int s_cnt, s_warp[EC_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cls = blockIdx.x;
    const int p3 = cls % P.sz, p2 = (cls / P.sz) % P.sy, p1 = cls / (P.sz * P.sy);
    // my pairs: b = 2 q, 2 q + 1; b' = 4 g .. 4 g + 3; my k vectors: slot, slot + ks, ... of every sub-chunk
    const int groups = nbp / 4, n_items = ((nb + 1) / 2) * groups;
    const int slot = tid / ipad, local = tid - slot * ipad;
    const int item = blockIdx.y * ipad + local;
    const bool has_item = slot < ks && item < n_items;
    const int b0 = has_item ? 2 * (item / groups) : 0, g4 = has_item ? (item % groups) * 4 : 0;
    double accC[2][4], accS[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) accC[j][i] = accS[j][i] = 0.0;
    for (int e = tid; e < 2 * EC_KCH * nbp; e += EC_THREADS) sm[e] = 0.0;

    // candidates of the class: n_i = p_i + s_i m_i inside [-kmax_i, kmax_i] (n3 >= 0: half space
    // n3 > 0 | (n3 == 0, n2 > 0) | (n3 == n2 == 0, n1 > 0), the one ewald.cu's k list uses)
    auto lo_m = [](int p, int s, int kmax) { return -((kmax + p) / s); };                // smallest m with n >= -kmax
    auto hi_m = [](int p, int s, int kmax) { return (kmax - p) >= 0 ? (kmax - p) / s : -1; };
    const int m1lo = lo_m(p1, P.sx, P.kmax[0]), m1hi = hi_m(p1, P.sx, P.kmax[0]);
    const int m2lo = lo_m(p2, P.sy, P.kmax[1]), m2hi = hi_m(p2, P.sy, P.kmax[1]);
    const int m3lo = 0, m3hi = hi_m(p3, P.sz, P.kmax[2]);
    const long long M1 = m1hi - m1lo + 1, M2 = m2hi - m2lo + 1, M3 = m3hi - m3lo + 1;
    const long long n_cand = (M1 > 0 && M2 > 0 && M3 > 0) ? M1 * M2 * M3 : 0;
    unsigned long long my_valid = 0;

    for (long long base = 0; base < n_cand; base += EC_THREADS) {
        // ---- test one candidate per thread, compact the valid ones into s_n / s_w
        const long long c = base + tid;
        bool ok = false;
        int n1 = 0, n2 = 0, n3 = 0;
        double w = 0.0;
        if (c < n_cand) {
            const int i3 = (int)(c % M3), i2 = (int)((c / M3) % M2), i1 = (int)(c / (M3 * M2));
            n1 = p1 + P.sx * (m1lo + i1);
            n2 = p2 + P.sy * (m2lo + i2);
            n3 = p3 + P.sz * (m3lo + i3);
            const bool half = n3 > 0 || (n3 == 0 && (n2 > 0 || (n2 == 0 && n1 > 0)));
            const double kx = n1 * P.B[0] + n2 * P.B[3] + n3 * P.B[6];
            const double ky = n1 * P.B[1] + n2 * P.B[4] + n3 * P.B[7];
            const double kz = n1 * P.B[2] + n2 * P.B[5] + n3 * P.B[8];
            const double k2 = kx * kx + ky * ky + kz * kz;
            ok = half && (k2 < P.kc2);
            if (ok) w = P.coeff * exp(-k2 / P.alpha4) / k2;   // core.py:869-875
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
        for (int q = 0; q < EC_THREADS / 32; ++q) {
            if (q < wid) off += s_warp[q];
            total += s_warp[q];
        }
        if (ok) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            s_n[3 * pos] = n1; s_n[3 * pos + 1] = n2; s_n[3 * pos + 2] = n3;
            s_w[pos] = w;
        }
        if (tid == 0) s_cnt = total;
        my_valid += ok ? 1ull : 0ull;
        __syncthreads();
        const int n_valid = s_cnt;
        // ---- accumulate the valid k of the batch, EC_KCH at a time
        for (int k0 = 0; k0 < n_valid; k0 += EC_KCH) {
            const int kn = min(EC_KCH, n_valid - k0);
            for (int e = tid; e < kn * nb; e += EC_THREADS) {
                const int kk = e / nb, bb = e - kk * nb;
                const int *nn = s_n + 3 * (k0 + kk);
                const double arg = fma((double)nn[0], theta[3 * bb], fma((double)nn[1], theta[3 * bb + 1],
                                                                          (double)nn[2] * theta[3 * bb + 2]));
                double sv, cv;
                sincos(arg, &sv, &cv);
                s_c[kk * nbp + bb] = cv;
                s_s[kk * nbp + bb] = sv;
            }
            __syncthreads();
            if (has_item) {
                for (int kk = slot; kk < kn; kk += ks) {
                    const double wk = s_w[k0 + kk];
                    const double *rc = s_c + kk * nbp, *rs = s_s + kk * nbp;
                    const double2 cb = *reinterpret_cast<const double2 *>(rc + b0);
                    const double2 sb = *reinterpret_cast<const double2 *>(rs + b0);
                    const double2 c01 = *reinterpret_cast<const double2 *>(rc + g4);
                    const double2 c23 = *reinterpret_cast<const double2 *>(rc + g4 + 2);
                    const double2 s01 = *reinterpret_cast<const double2 *>(rs + g4);
                    const double2 s23 = *reinterpret_cast<const double2 *>(rs + g4 + 2);
                    const double wc[2] = {wk * cb.x, wk * cb.y}, ws[2] = {wk * sb.x, wk * sb.y};
                    const double cp[4] = {c01.x, c01.y, c23.x, c23.y}, sp[4] = {s01.x, s01.y, s23.x, s23.y};
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            accC[j][i] = fma(wc[j], cp[i], fma(ws[j], sp[i], accC[j][i]));    // w cos(k.(tau_b' - tau_b))
                            accS[j][i] = fma(wc[j], sp[i], fma(-ws[j], cp[i], accS[j][i]));   // w sin(k.(tau_b' - tau_b))
                        }
                }
            }
            __syncthreads();
        }
    }
    // ---- partial sums of the k slots, added in slot order (the scratch aliases the panels: all reads are done)
    if (ks > 1) {
        __syncthreads();
        double *red = sm;                   // [slot][16][ipad]
        if (has_item && slot > 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    red[((slot * 16) + j * 4 + i) * ipad + local] = accC[j][i];
                    red[((slot * 16) + 8 + j * 4 + i) * ipad + local] = accS[j][i];
                }
        }
        __syncthreads();
        if (has_item && slot == 0)
            for (int q = 1; q < ks; ++q)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        accC[j][i] += red[((q * 16) + j * 4 + i) * ipad + local];
                        accS[j][i] += red[((q * 16) + 8 + j * 4 + i) * ipad + local];
                    }
    }
    if (has_item && slot == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (b0 + j >= nb) continue;
            const long long o = ((long long)cls * nb + b0 + j) * nb;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (g4 + i < nb) {
                    Cs[o + g4 + i] = accC[j][i];
                    Ss[o + g4 + i] = accS[j][i];
                }
        }
    }
    if (blockIdx.y == 0 && my_valid) atomicAdd(k_count, my_valid);
}

// ---- kernel B: transform over the classes.  f[b][(R, b')] = sum_p cos(phi_p(R)) C_p - sin(phi_p(R)) S_p.
// One thread per output element; phase factors from per-axis tables e^{2 pi i t / s_i} in shared memory.
__global__ void __launch_bounds__(256)
ewald_class_dft_kernel(CellParams P, const double *__restrict__ Cs, const double *__restrict__ Ss,
                       long long n_sites, double *__restrict__ f)
{
    extern __shared__ double tab[];   // cos / sin tables of the three axes
    const int nb = P.nb;
    double *cx = tab, *sxp = cx + P.sx, *cy = sxp + P.sx, *syp = cy + P.sy, *cz = syp + P.sy, *szp = cz + P.sz;
    for (int t = threadIdx.x; t < P.sx; t += blockDim.x) sincospi(2.0 * t / P.sx, &sxp[t], &cx[t]);
    for (int t = threadIdx.x; t < P.sy; t += blockDim.x) sincospi(2.0 * t / P.sy, &syp[t], &cy[t]);
    for (int t = threadIdx.x; t < P.sz; t += blockDim.x) sincospi(2.0 * t / P.sz, &szp[t], &cz[t]);
    __syncthreads();
    const long long total = (long long)nb * n_sites;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int b = (int)(idx / n_sites);
    const int j = (int)(idx - (long long)b * n_sites);
    const int cell = j / nb, bp = j - cell * nb;
    const int R3 = cell % P.sz, R2 = (cell / P.sz) % P.sy, R1 = cell / (P.sz * P.sy);
    const long long pair = (long long)b * nb + bp, stride = (long long)nb * nb;
    double acc = 0.0;
    int t1 = 0;
    for (int p1 = 0; p1 < P.sx; ++p1) {
        const double c1 = cx[t1], s1 = sxp[t1];
        int t2 = 0;
        for (int p2 = 0; p2 < P.sy; ++p2) {
            const double c12 = c1 * cy[t2] - s1 * syp[t2], s12 = s1 * cy[t2] + c1 * syp[t2];
            int t3 = 0;
            const long long cls0 = ((long long)p1 * P.sy + p2) * P.sz;
            for (int p3 = 0; p3 < P.sz; ++p3) {
                const double cph = c12 * cz[t3] - s12 * szp[t3], sph = s12 * cz[t3] + c12 * szp[t3];
                const long long o = (cls0 + p3) * stride + pair;
                acc = fma(cph, Cs[o], fma(-sph, Ss[o], acc));
                t3 += R3; t3 -= t3 >= P.sz ? P.sz : 0;
            }
            t2 += R2; t2 -= t2 >= P.sy ? P.sy : 0;
        }
        t1 += R1; t1 -= t1 >= P.sx ? P.sx : 0;
    }
    f[idx] = acc;
}

// ---- kernel B, separable form: the transform over the classes axis by axis through shared memory.  A CTA owns
// basis site b and PB consecutive b' (contiguous output runs): Z_p = C_p + i S_p of its pairs for ALL classes is
// loaded once (n_cells * PB complex numbers), then three passes (z, y, x) of s_axis complex multiply-adds per
// element -- n_cells (sx + sy + sz) instead of n_cells^2 operations per pair, and C_p, S_p are read exactly once.
template <int PB>
__global__ void __launch_bounds__(256)
ewald_class_dft3_kernel(CellParams P, const double *__restrict__ Cs, const double *__restrict__ Ss,
                        long long n_sites, double *__restrict__ f)
{
    extern __shared__ __align__(16) double sm3[];
    const int nb = P.nb, sx = P.sx, sy = P.sy, sz = P.sz;
    const int cells = sx * sy * sz;
    double2 *za = reinterpret_cast<double2 *>(sm3);            // [cells][PB]
    double2 *zb = za + (size_t)cells * PB;                     // [cells][PB]
    double *cx = reinterpret_cast<double *>(zb + (size_t)cells * PB);
    double *sxp = cx + sx, *cy = sxp + sx, *syp = cy + sy, *cz = syp + sy, *szp = cz + sz;
    const int tid = threadIdx.x;
    for (int t = tid; t < sx; t += blockDim.x) sincospi(2.0 * t / sx, &sxp[t], &cx[t]);
    for (int t = tid; t < sy; t += blockDim.x) sincospi(2.0 * t / sy, &syp[t], &cy[t]);
    for (int t = tid; t < sz; t += blockDim.x) sincospi(2.0 * t / sz, &szp[t], &cz[t]);
    const int b = blockIdx.x, bp0 = blockIdx.y * PB;
    const int total = cells * PB;
    for (int o = tid; o < total; o += blockDim.x) {
        const int q = o % PB, cls = o / PB;
        const int bp = min(bp0 + q, nb - 1);
        const long long src = ((long long)cls * nb + b) * nb + bp;
        za[o] = make_double2(Cs[src], Ss[src]);
    }
    __syncthreads();
    // pass z: zb[(p1,p2,R3)] = sum_p3 e^{2 pi i p3 R3 / sz} za[(p1,p2,p3)]
    for (int o = tid; o < total; o += blockDim.x) {
        const int q = o % PB, u = o / PB;
        const int R3 = u % sz, p12 = u / sz;
        double re = 0.0, im = 0.0;
        int t = 0;
        for (int p3 = 0; p3 < sz; ++p3) {
            const double2 z = za[(p12 * sz + p3) * PB + q];
            const double c = cz[t], sn = szp[t];
            re = fma(c, z.x, fma(-sn, z.y, re));
            im = fma(c, z.y, fma(sn, z.x, im));
            t += R3; t -= t >= sz ? sz : 0;
        }
        zb[o] = make_double2(re, im);
    }
    __syncthreads();
    // pass y: za[(p1,R2,R3)] = sum_p2 e^{2 pi i p2 R2 / sy} zb[(p1,p2,R3)]
    for (int o = tid; o < total; o += blockDim.x) {
        const int q = o % PB, u = o / PB;
        const int R3 = u % sz, R2 = (u / sz) % sy, p1 = u / (sz * sy);
        double re = 0.0, im = 0.0;
        int t = 0;
        for (int p2 = 0; p2 < sy; ++p2) {
            const double2 z = zb[((p1 * sy + p2) * sz + R3) * PB + q];
            const double c = cy[t], sn = syp[t];
            re = fma(c, z.x, fma(-sn, z.y, re));
            im = fma(c, z.y, fma(sn, z.x, im));
            t += R2; t -= t >= sy ? sy : 0;
        }
        za[o] = make_double2(re, im);
    }
    __syncthreads();
    // pass x (real part only): f[b][(R, b')] = Re sum_p1 e^{2 pi i p1 R1 / sx} za[(p1,R2,R3)]
    for (int o = tid; o < total; o += blockDim.x) {
        const int q = o % PB, u = o / PB;
        const int R23 = u % (sz * sy), R1 = u / (sz * sy);
        double re = 0.0;
        int t = 0;
        for (int p1 = 0; p1 < sx; ++p1) {
            const double2 z = za[(p1 * sy * sz + R23) * PB + q];
            re = fma(cx[t], z.x, fma(-sxp[t], z.y, re));
            t += R1; t -= t >= sx ? sx : 0;
        }
        if (bp0 + q < nb) f[(long long)b * n_sites + (long long)u * nb + bp0 + q] = re;
    }
}

}  // namespace pycd

using namespace pycd;

extern "C" int pycd_ewald_unit_rows(pycd_ctx *ctx, const pycd_ewald_desc *desc, int32_t n_basis,
                                    const int32_t size[3], double *out, pycd_ewald_stats *stats) {
    return guarded([&] {
        NvtxRange nvtx("pycd.ewald_unit_rows");
        PYCD_REQUIRE(ctx && desc && size && out, "NULL argument");
        const int nb = n_basis, sx = size[0], sy = size[1], sz = size[2];
        PYCD_REQUIRE(nb > 0 && nb <= 255 && sx > 0 && sy > 0 && sz > 0, "bad supercell");
        const long long cells = (long long)sx * sy * sz, n = cells * nb;
        PYCD_REQUIRE(n == desc->n_sites, "n_basis * cells != n_sites");
        PYCD_REQUIRE(desc->pbc[0] && desc->pbc[1] && desc->pbc[2], "the class factorisation needs pbc = [1, 1, 1]");
        PYCD_REQUIRE(desc->alpha > 0 && desc->volume > 0 && desc->dielectric > 0, "bad Ewald parameters");
        DeviceGuard g(ctx);
        cudaStream_t s = ctx->stream;
        // ---- what the factorisation assumes, verified on the host: sites = tau_b + x a1 + y a2 + z a3 with
        // a_i = (row i of the simulation cell) / s_i, and B reciprocal to that cell (the reference
        // builds the two matrices differently; they agree for the shapes it supports)
        std::vector<double> xyz((size_t)n * 3);
        PYCD_CUDA(cudaMemcpy(xyz.data(), desc->coords, sizeof(double) * n * 3, cudaMemcpyDefault));
        double a[3][3];
        const int ss[3] = {sx, sy, sz};
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) a[i][c] = desc->cell[3 * i + c] / ss[i];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double dot = desc->recip[3 * j] * a[i][0] + desc->recip[3 * j + 1] * a[i][1] + desc->recip[3 * j + 2] * a[i][2];
                const double want = (i == j) ? 2 * M_PI / ss[i] : 0.0;
                if (fabs(dot - want) > 1e-9)
                    throw Error("not translation-invariant: the reciprocal matrix is not the dual of the simulation cell");
            }
        double scale = 0.0;
        for (int c = 0; c < 9; ++c) scale = std::max(scale, fabs(desc->cell[c]));
        for (long long cell = 0; cell < cells; ++cell) {
            const int z = (int)(cell % sz), y = (int)((cell / sz) % sy), x = (int)(cell / ((long long)sz * sy));
            for (int b = 0; b < nb; ++b)
                for (int c = 0; c < 3; ++c) {
                    const double want = xyz[3 * b + c] + x * a[0][c] + y * a[1][c] + z * a[2][c];
                    if (fabs(xyz[3 * (cell * nb + b) + c] - want) > 1e-9 * scale)
                        throw Error("not translation-invariant: site coordinates are not unit cell 0 plus lattice translations");
                }
        }
        CellParams P;
        P.nb = nb; P.sx = sx; P.sy = sy; P.sz = sz;
        for (int c = 0; c < 3; ++c) P.kmax[c] = desc->k_max[c];
        for (int c = 0; c < 9; ++c) P.B[c] = desc->recip[c];
        P.coeff = (2 * M_PI) / desc->volume;
        P.alpha4 = 4 * desc->alpha;
        P.kc2 = desc->k_cut * desc->k_cut;
        std::vector<double> theta_h((size_t)nb * 3);
        for (int b = 0; b < nb; ++b)
            for (int j = 0; j < 3; ++j)
                theta_h[3 * b + j] = desc->recip[3 * j] * xyz[3 * b] + desc->recip[3 * j + 1] * xyz[3 * b + 1] +
                                     desc->recip[3 * j + 2] * xyz[3 * b + 2];
        // work buffers from the context's arena (six cudaMalloc / cudaFree pairs cost more than the kernels)
        struct { double *p; } theta, Cs, Ss, f, coords_dev;
        struct { unsigned long long *p; } kcount;
        const size_t cs_n = (size_t)cells * nb * nb;
        Arena arena(ctx);
        arena.reserve(sizeof(double) * theta_h.size());
        arena.reserve(sizeof(double) * (size_t)n * 3);
        arena.reserve(sizeof(double) * cs_n);
        arena.reserve(sizeof(double) * cs_n);
        arena.reserve(sizeof(double) * (size_t)nb * n);
        arena.reserve(sizeof(unsigned long long));
        arena.commit();
        theta.p = arena.take<double>(theta_h.size());
        coords_dev.p = arena.take<double>((size_t)n * 3);
        Cs.p = arena.take<double>(cs_n);
        Ss.p = arena.take<double>(cs_n);
        f.p = arena.take<double>((size_t)nb * n);
        kcount.p = arena.take<unsigned long long>(1);
        PYCD_CUDA(cudaMemcpyAsync(theta.p, theta_h.data(), sizeof(double) * theta_h.size(), cudaMemcpyHostToDevice, s));
        if (is_device_pointer(desc->coords)) coords_dev.p = const_cast<double *>(desc->coords);
        else PYCD_CUDA(cudaMemcpyAsync(coords_dev.p, xyz.data(), sizeof(double) * n * 3, cudaMemcpyHostToDevice, s));
        PYCD_CUDA(cudaMemsetAsync(kcount.p, 0, sizeof(unsigned long long), s));
        OutBuf<double> o;
        o.bind(out, (size_t)nb * n);

        KernelTimer tf(ctx, KC_EWALD_FOURIER);
        {
            const int nbp = (nb + 3) & ~3;
            const int n_items = ((nb + 1) / 2) * (nbp / 4);
            const int ipad = std::min(EC_THREADS, (n_items + 31) & ~31);
            const int ks = std::min(EC_KCH, EC_THREADS / ipad);
            const unsigned tiles = (unsigned)((n_items + ipad - 1) / ipad);
            const size_t panels = sizeof(double) * (2 * (size_t)EC_KCH * nbp + EC_THREADS) + sizeof(int) * 3 * EC_THREADS;
            const size_t smem = std::max(panels, ks > 1 ? sizeof(double) * 16 * (size_t)ks * ipad : (size_t)0);
            PYCD_REQUIRE(cells < (1ll << 31) && tiles < 65536, "grid too large");
            ewald_class_sums_kernel<<<dim3((unsigned)cells, tiles), EC_THREADS, smem, s>>>(P, theta.p, Cs.p, Ss.p, kcount.p,
                                                                                           ipad, ks);
            check_launch(ctx, "ewald_class_sums_kernel");
            // separable transform when the classes of 4 / 2 / 1 pairs fit in shared memory twice, else the direct one
            const size_t tsm = sizeof(double) * 2 * (size_t)(sx + sy + sz);
            auto need = [&](int pb) { return 2 * sizeof(double2) * (size_t)cells * pb + tsm; };
            const size_t cap = 200 * 1024;
            const char *force = getenv("PYCD_EWALD_DFT");   // "direct": A/B against the one-thread-per-element form
            const bool direct = force && std::string(force) == "direct";
            if (!direct && need(4) <= cap) {
                PYCD_CUDA(cudaFuncSetAttribute(ewald_class_dft3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(4)));
                ewald_class_dft3_kernel<4><<<dim3(nb, (nb + 3) / 4), 256, need(4), s>>>(P, Cs.p, Ss.p, n, f.p);
            } else if (!direct && need(2) <= cap) {
                PYCD_CUDA(cudaFuncSetAttribute(ewald_class_dft3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(2)));
                ewald_class_dft3_kernel<2><<<dim3(nb, (nb + 1) / 2), 256, need(2), s>>>(P, Cs.p, Ss.p, n, f.p);
            } else if (!direct && need(1) <= cap) {
                PYCD_CUDA(cudaFuncSetAttribute(ewald_class_dft3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(1)));
                ewald_class_dft3_kernel<1><<<dim3(nb, nb), 256, need(1), s>>>(P, Cs.p, Ss.p, n, f.p);
            } else {
                const long long total = (long long)nb * n;
                ewald_class_dft_kernel<<<(unsigned)((total + 255) / 256), 256, tsm, s>>>(P, Cs.p, Ss.p, n, f.p);
            }
            check_launch(ctx, "ewald_class_dft_kernel");
        }
        tf.stop(2);
        KernelTimer tr(ctx, KC_EWALD_FINISH);
        ewald_finish_rows(ctx, desc, coords_dev.p, n, nb, f.p, o.dev());
        tr.stop(1);
        unsigned long long k_eff = 0;
        PYCD_CUDA(cudaMemcpyAsync(&k_eff, kcount.p, sizeof(k_eff), cudaMemcpyDeviceToHost, s));
        o.finish(s);
        PYCD_CUDA(cudaStreamSynchronize(s));
        tf.read();
        tr.read();
        if (stats) {
            stats->k_eff = (int64_t)k_eff;
            stats->rows = nb;
            stats->fourier_ms = ctx->last_ms[KC_EWALD_FOURIER];
            stats->finish_ms = ctx->last_ms[KC_EWALD_FINISH];
            stats->flops = 8.0 * (double)nb * nb * (double)k_eff + 4.0 * (double)nb * (double)n * (double)cells +
                           40.0 * (double)nb * (double)n;
            stats->k_split = 1;
        }
    });
}
