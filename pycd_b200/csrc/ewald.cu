// ewald.cu -- fp64 Ewald site-pair array on sm_100a.
//
// P[i,j] = R_ij/eps + F_ij/eps - delta_ij*sqrt(alpha/pi)/eps      (PyCD/core.py:1659-1661)
//   R_ij = erfc(sqrt(alpha) r_ij)/(2 r_ij)  for r_ij < r_cut, R_ii = 0.5   (core.py:799-814)
//   F_ij = 2 * sum_{k in half space, |k|^2<k_cut^2} (2pi/V) exp(-k^2/4alpha) cos(k.d_ij)/k^2
//                                                                        (core.py:853-878)
// The reference evaluates cos(k.d_ij) per pair per k from an (N,N,3) array.  Here
// cos(k.(r_j-r_i)) = c_i c_j + s_i s_j turns the k sum into a rank-2K update
// F = A A^T with A[i,(k,0/1)] = sqrt-free weighted (cos,sin)(k.r_i): a GEMM whose
// operand panels are generated on the fly in shared memory (A would be 403 GB at
// N=30000), accumulated with DFMA in registers.  fp64 has no tcgen05 path; the
// bound is the FP64 pipe (64 DFMA/clk/SM).
#include "common.cuh"

#include <cmath>
#include <cstdint>
#include <type_traits>
#include <cstdlib>
#include <algorithm>

namespace pycd {

// one half-space k vector: integer triple + flag bit0 = "continues the previous
// entry along n3" (phase obtained by one complex multiply instead of a sincos)
struct __align__(8) KEntry {
    short n1, n2, n3, flag;
};

constexpr int EW_THREADS = 256;

template <int BM, int BN, int TM, int TN, int KC>
struct EwaldTile {
    static constexpr int TX = BN / TN;
    static constexpr int TY = BM / TM;
    static_assert(TX * TY == EW_THREADS, "tile/thread shape mismatch");
    static_assert(TM % 2 == 0 && TN % 2 == 0, "double2 operand loads");
    static constexpr size_t smem_bytes =
        sizeof(double) * (size_t)(2 * KC) * (BM + BN) + sizeof(KEntry) * KC + sizeof(double) * KC;
};

// theta[j][a] = B_a . r_j  so that k_n . r_j = n1*theta_j0 + n2*theta_j1 + n3*theta_j2
__global__ void ewald_theta_kernel(const double *__restrict__ coords, long long n,
                                   double b00, double b01, double b02, double b10, double b11,
                                   double b12, double b20, double b21, double b22,
                                   double *__restrict__ theta)
{
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= n) return;
    double x = coords[3 * j], y = coords[3 * j + 1], z = coords[3 * j + 2];
    theta[3 * j + 0] = b00 * x + b01 * y + b02 * z;
    theta[3 * j + 1] = b10 * x + b11 * y + b12 * z;
    theta[3 * j + 2] = b20 * x + b21 * y + b22 * z;
}

// (cos, sin)(k.r) of one site along a run of the k list: a sincos where a run starts, one
// complex multiply by e^{i theta_3} where the entry continues the previous one along n3.
// start of a run of the k list: full-accuracy sincos of n.theta.  Kept out of line: it is taken for
// ~1 % of the entries and its range reduction would otherwise be inlined into every unrolled step.
__device__ __noinline__ void phase_restart(double t1, double t2, double t3, int n1, int n2, int n3,
                                           double *c, double *s)
{
    const double arg = fma((double)n1, t1, fma((double)n2, t2, (double)n3 * t3));
    sincos(arg, s, c);
}

struct PhaseWalker {
    double t1, t2, t3, c3, s3, c, s;
    __device__ __forceinline__ void init(const double *__restrict__ theta, long long site) {
        t1 = theta[3 * site]; t2 = theta[3 * site + 1]; t3 = theta[3 * site + 2];
        sincos(t3, &s3, &c3);
        c = 1.0; s = 0.0;
    }
    __device__ __forceinline__ void step(const KEntry ke, bool restart) {
        if (restart || !(ke.flag & 1)) {
            const double arg = fma((double)ke.n1, t1, fma((double)ke.n2, t2, (double)ke.n3 * t3));
            sincos(arg, &s, &c);
        } else {
            const double cn = c * c3 - s * s3;
            s = fma(s, c3, c * s3);
            c = cn;
        }
    }
    // same, with the restart out of line (pipelined kernel: the step sits inside the DMMA loop)
    __device__ __forceinline__ void step_lean(const KEntry ke, bool restart) {
        if (restart || !(ke.flag & 1)) {
            double cc, ss;
            phase_restart(t1, t2, t3, ke.n1, ke.n2, ke.n3, &cc, &ss);
            c = cc; s = ss;
        } else {
            const double cn = c * c3 - s * s3;
            s = fma(s, c3, c * s3);
            c = cn;
        }
    }
};

// Walker of the pipelined kernel: only the recurrence state lives in registers; a run start (rare)
// re-reads the site's theta from global memory and calls the out-of-line sincos.
struct LeanWalker {
    double c3, s3, c, s;
    const double *th;
    __device__ __forceinline__ void init(const double *__restrict__ theta, long long site) {
        th = theta + 3 * site;
        sincos(th[2], &s3, &c3);
        c = 1.0; s = 0.0;
    }
    __device__ __forceinline__ void restart(int n1, int n2, int n3) {
        double cc, ss;
        phase_restart(th[0], th[1], th[2], n1, n2, n3, &cc, &ss);
        c = cc; s = ss;
    }
    __device__ __forceinline__ void step_lean(const KEntry ke, bool force) {
        if (force || !(ke.flag & 1)) {
            restart(ke.n1, ke.n2, ke.n3);
        } else {
            const double cn = c * c3 - s * s3;
            s = fma(s, c3, c * s3);
            c = cn;
        }
    }
};

// Reciprocal-space partial sums.  grid = k_split * n_row_tiles * n_col_tiles CTAs of
// 256 threads; CTA (ks, rt, ct) accumulates its share of the k chunks for the
// BM x BN tile and writes out[ks][row][col] (un-scaled: sum_k w_k cos(k.d_ij)).
// Panel generation: thread t walks tile site t (rows first, then columns) through every
// entry of the chunk, carrying its phase across chunks; when the tile has more sites than
// threads (skinny tile: BN = 256 columns + BM rows) the BM row sites are walked in
// KC/ (256/BM)-entry pieces by all threads.  CTAS CTAs share an SM so that one CTA's panel
// generation overlaps another's DFMA phase.
template <int BM, int BN, int TM, int TN, int KC, int CTAS>
__global__ void __launch_bounds__(EW_THREADS, CTAS)
ewald_fourier_kernel(const double *__restrict__ theta, long long n_sites, long long row0,
                     long long n_rows, const KEntry *__restrict__ kent,
                     const double *__restrict__ kw, int n_chunks, int k_split, int n_row_tiles,
                     int n_col_tiles, double *__restrict__ out)
{
    using T = EwaldTile<BM, BN, TM, TN, KC>;
    constexpr int TX = T::TX, TY = T::TY;
    constexpr bool SPLIT_ROWS = (BM + BN > EW_THREADS);   // skinny tile
    static_assert(!SPLIT_ROWS || BN == EW_THREADS, "skinny tile: one column site per thread");
    static_assert(SPLIT_ROWS || BM + BN == EW_THREADS, "dense tile: one tile site per thread");
    constexpr int ROW_PIECES = SPLIT_ROWS ? EW_THREADS / BM : 1;
    constexpr int PIECE = KC / ROW_PIECES;
    static_assert(!SPLIT_ROWS || (KC % ROW_PIECES == 0 && PIECE >= 1), "row pieces must tile the chunk");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);  // [2*KC][BM]  rows, weighted
    double *Bs = As + 2 * KC * BM;                       // [2*KC][BN]  columns
    double *s_w = Bs + 2 * KC * BN;                      // [KC]
    KEntry *s_ent = reinterpret_cast<KEntry *>(s_w + KC);  // [KC]

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    int bid = blockIdx.x;
    const int ct = bid % n_col_tiles; bid /= n_col_tiles;
    const int rt = bid % n_row_tiles; bid /= n_row_tiles;
    const int ks = bid;
    const int c0 = (int)(((long long)ks * n_chunks) / k_split);
    const int c1 = (int)(((long long)(ks + 1) * n_chunks) / k_split);

    // tile-site -> global site (clamped; out-of-range results are never stored)
    auto row_site = [&](int r) -> long long {
        const long long s = row0 + (long long)rt * BM + r;
        return s < n_sites ? s : n_sites - 1;
    };
    auto col_site = [&](int c) -> long long {
        const long long s = (long long)ct * BN + c;
        return s < n_sites ? s : n_sites - 1;
    };
    // primary walker: dense tile -> tile site tid (rows then columns); skinny -> column tid
    const bool prim_is_row = !SPLIT_ROWS && tid < BM;
    const int prim_idx = SPLIT_ROWS ? tid : (prim_is_row ? tid : tid - BM);
    PhaseWalker pw, rw;
    pw.init(theta, prim_is_row ? row_site(prim_idx) : col_site(prim_idx));
    const int r_site = tid % BM, r_piece = tid / BM;      // skinny: row walker assignment
    if (SPLIT_ROWS) rw.init(theta, row_site(r_site));

    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;

    for (int chunk = c0; chunk < c1; ++chunk) {
        __syncthreads();  // previous product finished with the panels
        if (tid < KC) {
            s_ent[tid] = kent[(long long)chunk * KC + tid];
            s_w[tid] = kw[(long long)chunk * KC + tid];
        }
        __syncthreads();
        // ---- generate the operand panels ----
#pragma unroll 4
        for (int e = 0; e < KC; ++e) {
            pw.step(s_ent[e], chunk == c0 && e == 0);
            if (prim_is_row) {
                const double w = s_w[e];
                As[e * BM + prim_idx] = w * pw.c;
                As[(KC + e) * BM + prim_idx] = w * pw.s;
            } else {
                Bs[e * BN + prim_idx] = pw.c;
                Bs[(KC + e) * BN + prim_idx] = pw.s;
            }
        }
        if (SPLIT_ROWS) {
#pragma unroll
            for (int q = 0; q < PIECE; ++q) {
                const int e = r_piece * PIECE + q;
                rw.step(s_ent[e], q == 0);
                const double w = s_w[e];
                As[e * BM + r_site] = w * rw.c;
                As[(KC + e) * BM + r_site] = w * rw.s;
            }
        }
        __syncthreads();
        // ---- rank-2KC update of the register tile, operands double-buffered in registers ----
        double a[2][TM], b[2][TN];
        auto load_ab = [&](int buf, int kk) {
#pragma unroll
            for (int i = 0; i < TM / 2; ++i) {
                const double2 v = *reinterpret_cast<const double2 *>(&As[kk * BM + i * 2 * TY + 2 * ty]);
                a[buf][2 * i] = v.x;
                a[buf][2 * i + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < TN / 2; ++j) {
                const double2 v = *reinterpret_cast<const double2 *>(&Bs[kk * BN + j * 2 * TX + 2 * tx]);
                b[buf][2 * j] = v.x;
                b[buf][2 * j + 1] = v.y;
            }
        };
        load_ab(0, 0);
        constexpr int UNR = (TM * TN >= 64) ? 8 : 2 * KC;   // keep the loop body inside the i-cache
#pragma unroll(UNR)
        for (int kk = 0; kk < 2 * KC; ++kk) {
            if (kk + 1 < 2 * KC) load_ab((kk + 1) & 1, kk + 1);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[kk & 1][i], b[kk & 1][j], acc[i][j]);
        }
    }

    double *o = out + (long long)ks * n_rows * n_sites;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long row = (long long)rt * BM + (i / 2) * 2 * TY + 2 * ty + (i & 1);
        if (row >= n_rows) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const long long col = (long long)ct * BN + (j / 2) * 2 * TX + 2 * tx + (j & 1);
            if (col < n_sites) o[row * n_sites + col] = acc[i][j];
        }
    }
}

// mma.sync.m8n8k4.f64: the fp64 tensor-core path of sm_100 (tcgen05 has no fp64 kind); 256
// FMAs per warp instruction.  Panels are stored as [k/4][site][4] so that the A (row) and
// B (col) fragments of a k4 slice are one conflict-free LDS.64 per lane (lane l reads element
// [site0 + l/4][k0 + l%4] = base + l).
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------
// Phased DMMA kernel (the default): all 8 warps generate the panels of a chunk, then all 8
// accumulate with mma.sync.m8n8k4.f64; two CTAs share an SM so that one CTA's generation
// phase runs under the other's DMMA stream.  (Dedicated producer warps starve: their
// dependent DFMA chains queue behind 16-cycle DMMA issues on the shared FP64 pipe.)
// Measured register-only ceilings on this B200 (tools/fp64_peak.cu): DMMA 37.1 TFLOP/s,
// register-tiled DFMA 25-28 TFLOP/s.
template <int BM, int BN, int WM, int WN, int KC>
__global__ void __launch_bounds__(EW_THREADS, 2)
ewald_fourier_dmma_kernel(const double *__restrict__ theta, long long n_sites, long long row0,
                          long long n_rows, const KEntry *__restrict__ kent,
                          const double *__restrict__ kw, int n_chunks, int k_split, int n_row_tiles,
                          int n_col_tiles, double *__restrict__ out)
{
    static_assert((BM / WM) * (BN / WN) == EW_THREADS / 32, "8 warps must tile the CTA tile");
    static_assert(WM % 8 == 0 && WN % 8 == 0 && KC % 2 == 0, "m8n8k4 fragments");
    constexpr int NS = BM + BN;
    constexpr bool SPLIT_ROWS = (NS > EW_THREADS);        // skinny tile: BN == 256 columns + BM rows
    static_assert(!SPLIT_ROWS || BN == EW_THREADS, "skinny tile: one column site per thread");
    constexpr int ROW_PIECES = SPLIT_ROWS ? EW_THREADS / BM : 1;
    constexpr int PIECE = KC / ROW_PIECES;
    static_assert(!SPLIT_ROWS || (KC % ROW_PIECES == 0 && PIECE >= 1), "row pieces must tile the chunk");
    constexpr int MT = WM / 8, NT = WN / 8, KG = KC / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *As = reinterpret_cast<double *>(smem_raw);    // [KG][BM][4]  rows, weighted
    double *Bs = As + 2 * KC * BM;                         // [KG][BN][4]  columns
    double *s_w = Bs + 2 * KC * BN;                        // [KC]
    KEntry *s_ent = reinterpret_cast<KEntry *>(s_w + KC);  // [KC]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int bid = blockIdx.x;
    const int ct = bid % n_col_tiles; bid /= n_col_tiles;
    const int rt = bid % n_row_tiles; bid /= n_row_tiles;
    const int ks = bid;
    const int c0 = (int)(((long long)ks * n_chunks) / k_split);
    const int c1 = (int)(((long long)(ks + 1) * n_chunks) / k_split);

    auto row_site = [&](int r) -> long long {
        const long long s = row0 + (long long)rt * BM + r;
        return s < n_sites ? s : n_sites - 1;
    };
    auto col_site = [&](int c) -> long long {
        const long long s = (long long)ct * BN + c;
        return s < n_sites ? s : n_sites - 1;
    };
    // primary walker: dense tile -> tile site tid (rows first, then columns; threads >= NS
    // idle during generation); skinny -> column tid
    const bool prim_active = SPLIT_ROWS || tid < NS;
    const bool prim_is_row = !SPLIT_ROWS && tid < BM;
    const int prim_idx = SPLIT_ROWS ? tid : (prim_is_row ? tid : tid - BM);
    PhaseWalker pw, rw;
    pw.init(theta, prim_is_row ? row_site(prim_idx) : col_site(prim_active ? prim_idx : 0));
    const int r_site = tid % BM, r_piece = tid / BM;
    if (SPLIT_ROWS) rw.init(theta, row_site(r_site));

    const int m0 = (wid % (BM / WM)) * WM, n0 = (wid / (BM / WM)) * WN;
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int chunk = c0; chunk < c1; ++chunk) {
        __syncthreads();  // previous product finished with the panels
        if (tid < KC) {
            s_ent[tid] = kent[(long long)chunk * KC + tid];
            s_w[tid] = kw[(long long)chunk * KC + tid];
        }
        __syncthreads();
        if (prim_active) {
#pragma unroll 4
            for (int e = 0; e < KC; ++e) {
                pw.step(s_ent[e], chunk == c0 && e == 0);
                if (prim_is_row) {
                    const double w = s_w[e];
                    *reinterpret_cast<double2 *>(&As[((e >> 1) * BM + prim_idx) * 4 + (e & 1) * 2]) =
                        make_double2(w * pw.c, w * pw.s);
                } else {
                    *reinterpret_cast<double2 *>(&Bs[((e >> 1) * BN + prim_idx) * 4 + (e & 1) * 2]) =
                        make_double2(pw.c, pw.s);
                }
            }
        }
        if (SPLIT_ROWS) {
#pragma unroll
            for (int q = 0; q < PIECE; ++q) {
                const int e = r_piece * PIECE + q;
                rw.step(s_ent[e], q == 0);
                const double w = s_w[e];
                *reinterpret_cast<double2 *>(&As[((e >> 1) * BM + r_site) * 4 + (e & 1) * 2]) =
                    make_double2(w * rw.c, w * rw.s);
            }
        }
        __syncthreads();
        const double *Aw = As + (size_t)m0 * 4 + lane;
        const double *Bw = Bs + (size_t)n0 * 4 + lane;
#pragma unroll
        for (int g = 0; g < KG; ++g) {
            double a[MT], b[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) a[i] = Aw[(g * BM + i * 8) * 4];
#pragma unroll
            for (int j = 0; j < NT; ++j) b[j] = Bw[(g * BN + j * 8) * 4];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }

    double *o = out + (long long)ks * n_rows * n_sites;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const long long row = (long long)rt * BM + m0 + i * 8 + (lane >> 2);
        if (row >= n_rows) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const long long col = (long long)ct * BN + n0 + j * 8 + (lane & 3) * 2;
            if (col < n_sites) o[row * n_sites + col] = acc[i][j][0];
            if (col + 1 < n_sites) o[row * n_sites + col + 1] = acc[i][j][1];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Software-pipelined DMMA kernel: the panels are double-buffered and every warp generates the
// NEXT chunk's panel entries between the DMMA groups of the CURRENT chunk (two k entries per
// k4 slice), with one block barrier per chunk instead of two.  A chunk without a run start
// (~7 of 8 at cfg 3) takes a branch-free path -- the phase recurrence only, no k entries, no
// sincos -- so that the compiler can spread the scalar FP64 chain between the DMMAs; a chunk
// with a run start takes the generic path.  Skinny tile (32 x 256): thread t owns column t; the
// 32 row sites are covered by all 8 warps, warp w generating the row entries of k4 slice w (two
// entries) with a carried phase that jumps 15 entries ahead between chunks (e^{i 15 theta_3}).
// Measured context (tools/fp64_peak.cu): a 4x4 DMMA tile fed from shared memory reaches 36.9
// TFLOP/s; every complex-multiply step per 16 DMMAs issued beside it costs ~5 %.
template <int BM, int BN, int WM, int WN, int KC, int CTAS>
__global__ void __launch_bounds__(EW_THREADS, CTAS)
ewald_fourier_pipe_kernel(const double *__restrict__ theta, long long n_sites, long long row0,
                          long long n_rows, const KEntry *__restrict__ kent,
                          const double *__restrict__ kw, int n_chunks, int k_split, int n_row_tiles,
                          int n_col_tiles, double *__restrict__ out)
{
    static_assert((BM / WM) * (BN / WN) == EW_THREADS / 32, "8 warps must tile the CTA tile");
    static_assert(WM % 8 == 0 && WN % 8 == 0 && KC % 2 == 0 && KC <= 32, "m8n8k4 fragments");
    constexpr int NS = BM + BN;
    constexpr bool SKINNY = (NS > EW_THREADS);            // BN == 256 columns (one per thread) + 32 rows
    static_assert(!SKINNY || (BN == EW_THREADS && BM == 32 && KC == 16), "skinny tile: 32 x 256, 16-entry chunks");
    constexpr int MT = WM / 8, NT = WN / 8, KG = KC / 2;
    constexpr int PANEL = 2 * KC * NS;                    // doubles per buffer: As [KG][BM][4], then Bs [KG][BN][4]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *panels = reinterpret_cast<double *>(smem_raw);          // [2][PANEL]
    double *s_w = panels + 2 * PANEL;                                // [2][KC]
    KEntry *s_ent = reinterpret_cast<KEntry *>(s_w + 2 * KC);        // [2][KC]
    unsigned *s_mask = reinterpret_cast<unsigned *>(s_ent + 2 * KC); // [2] bit e: entry e starts a run

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int bid = blockIdx.x;
    const int ct = bid % n_col_tiles; bid /= n_col_tiles;
    const int rt = bid % n_row_tiles; bid /= n_row_tiles;
    const int ks = bid;
    const int c0 = (int)(((long long)ks * n_chunks) / k_split);
    const int c1 = (int)(((long long)(ks + 1) * n_chunks) / k_split);

    auto row_site = [&](int r) -> long long {
        const long long s = row0 + (long long)rt * BM + r;
        return s < n_sites ? s : n_sites - 1;
    };
    auto col_site = [&](int c) -> long long {
        const long long s = (long long)ct * BN + c;
        return s < n_sites ? s : n_sites - 1;
    };
    // primary walker: skinny -> column tid; dense -> tile site tid (rows first, then columns)
    const bool prim_active = SKINNY || tid < NS;
    const bool prim_is_row = !SKINNY && tid < BM;
    const int prim_idx = SKINNY ? tid : (prim_is_row ? tid : tid - BM);
    LeanWalker pw, rw;
    pw.init(theta, prim_is_row ? row_site(prim_idx) : col_site(prim_active ? prim_idx : 0));
    double c15 = 1.0, s15 = 0.0;                          // skinny: e^{i 15 theta_3} of row `lane`
    if (SKINNY) {
        rw.init(theta, row_site(lane));
        sincos(15.0 * rw.th[2], &s15, &c15);
    }

    // k entries + weights of a chunk travel through registers of warp 0 (loaded one iteration before
    // they are stored, so that the global-load latency hides under that iteration's DMMAs); the store
    // also publishes the chunk's run-start mask
    KEntry e_pre = {0, 0, 0, 0};
    double w_pre = 0.0;
    auto fetch_entries = [&](int chunk) {
        if (tid < KC && chunk < c1) {
            e_pre = kent[(long long)chunk * KC + tid];
            w_pre = kw[(long long)chunk * KC + tid];
        }
    };
    auto store_entries = [&](int b) {
        if (tid < 32) {
            if (tid < KC) {
                s_ent[b * KC + tid] = e_pre;
                s_w[b * KC + tid] = w_pre;
            }
            const unsigned m = __ballot_sync(0xffffffffu, tid < KC && !(e_pre.flag & 1));
            if (tid == 0) s_mask[b] = m;
        }
    };
    auto store4 = [](double *dst, double a0, double a1, double a2, double a3) {
        *reinterpret_cast<double2 *>(dst) = make_double2(a0, a1);
        *reinterpret_cast<double2 *>(dst + 2) = make_double2(a2, a3);
    };
    // generic path: panel entries of k4 slice g (k entries 2g, 2g+1), any run structure
    auto gen_pair = [&](double *buf, const KEntry *ent, const double *w, int g, bool first) {
        const KEntry e0 = ent[2 * g], e1 = ent[2 * g + 1];
        if (prim_active) {
            pw.step_lean(e0, first);
            double a0 = pw.c, a1 = pw.s;
            pw.step_lean(e1, false);
            double a2 = pw.c, a3 = pw.s;
            if (prim_is_row) {
                const double w0 = w[2 * g], w1 = w[2 * g + 1];
                store4(buf + (g * BM + prim_idx) * 4, a0 * w0, a1 * w0, a2 * w1, a3 * w1);
            } else {
                store4(buf + 2 * KC * BM + (g * BN + prim_idx) * 4, a0, a1, a2, a3);
            }
        }
        if (SKINNY && g == wid) {   // this warp's row slice: restart at its first entry
            const double w0 = w[2 * g], w1 = w[2 * g + 1];
            rw.step_lean(e0, true);
            const double a0 = w0 * rw.c, a1 = w0 * rw.s;
            rw.step_lean(e1, false);
            store4(buf + (g * BM + lane) * 4, a0, a1, w1 * rw.c, w1 * rw.s);
            // carried row phase = the chunk's LAST run taken back to this slice's second entry, so that
            // a following run-continuing chunk finds its entry 2g exactly 15 entries ahead
            const KEntry el = ent[KC - 1];
            rw.restart(el.n1, el.n2, el.n3 - (KC - 2 - 2 * g));
        }
    };
    // branch-free path for a chunk that continues every run: phase recurrence only
    auto gen_pair_fast = [&](double *buf, const double *w, int g) {
        if (SKINNY || prim_active) {
            const double a0 = pw.c * pw.c3 - pw.s * pw.s3;
            const double a1 = fma(pw.s, pw.c3, pw.c * pw.s3);
            const double a2 = a0 * pw.c3 - a1 * pw.s3;
            const double a3 = fma(a1, pw.c3, a0 * pw.s3);
            pw.c = a2; pw.s = a3;
            if (!SKINNY && prim_is_row) {
                const double w0 = w[2 * g], w1 = w[2 * g + 1];
                store4(buf + (g * BM + prim_idx) * 4, a0 * w0, a1 * w0, a2 * w1, a3 * w1);
            } else {
                store4(buf + 2 * KC * BM + (g * BN + prim_idx) * 4, a0, a1, a2, a3);
            }
        }
        if (SKINNY && g == wid) {   // rows: 15 entries ahead of the previous chunk's last one, then one more
            const double w0 = w[2 * g], w1 = w[2 * g + 1];
            const double a0 = rw.c * c15 - rw.s * s15;
            const double a1 = fma(rw.s, c15, rw.c * s15);
            const double a2 = a0 * rw.c3 - a1 * rw.s3;
            const double a3 = fma(a1, rw.c3, a0 * rw.s3);
            rw.c = a2; rw.s = a3;
            store4(buf + (g * BM + lane) * 4, a0 * w0, a1 * w0, a2 * w1, a3 * w1);
        }
    };

    const int m0 = (wid % (BM / WM)) * WM, n0 = (wid / (BM / WM)) * WN;
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue: k entries of the first two chunks, panel of the first (generic path, forced start)
    fetch_entries(c0);
    store_entries(0);
    fetch_entries(c0 + 1);
    store_entries(1);
    fetch_entries(c0 + 2);
    __syncthreads();
    if (c0 < c1) {
#pragma unroll
        for (int g = 0; g < KG; ++g) gen_pair(panels, s_ent, s_w, g, g == 0);
    }
    __syncthreads();

    for (int chunk = c0; chunk < c1; ++chunk) {
        const int cur = (chunk - c0) & 1, nxt = cur ^ 1;
        const bool has_next = chunk + 1 < c1;
        const bool fast = has_next && s_mask[nxt] == 0u;
        // k entries of chunk + 2 (in registers since the previous iteration) replace those of `chunk`,
        // consumed one iteration ago and read again only after this iteration's barrier
        if (chunk + 2 < c1) store_entries(cur);
        fetch_entries(chunk + 3);
        const double *Aw = panels + cur * PANEL + (size_t)m0 * 4 + lane;
        const double *Bw = panels + cur * PANEL + 2 * KC * BM + (size_t)n0 * 4 + lane;
        double *nbuf = panels + nxt * PANEL;
        const KEntry *nent = s_ent + nxt * KC;
        const double *nw = s_w + nxt * KC;
        if (fast) {
#pragma unroll
            for (int g = 0; g < KG; ++g) {
                double a[MT], b[NT];
#pragma unroll
                for (int i = 0; i < MT; ++i) a[i] = Aw[(g * BM + i * 8) * 4];
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = Bw[(g * BN + j * 8) * 4];
                gen_pair_fast(nbuf, nw, g);
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        } else {
#pragma unroll 1
            for (int g = 0; g < KG; ++g) {
                double a[MT], b[NT];
#pragma unroll
                for (int i = 0; i < MT; ++i) a[i] = Aw[(g * BM + i * 8) * 4];
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = Bw[(g * BN + j * 8) * 4];
                if (has_next) gen_pair(nbuf, nent, nw, g, false);
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        __syncthreads();   // next panel complete; this panel and the next chunk's k entries consumed
    }

    double *o = out + (long long)ks * n_rows * n_sites;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const long long row = (long long)rt * BM + m0 + i * 8 + (lane >> 2);
        if (row >= n_rows) continue;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const long long col = (long long)ct * BN + n0 + j * 8 + (lane & 3) * 2;
            if (col < n_sites) o[row * n_sites + col] = acc[i][j][0];
            if (col + 1 < n_sites) o[row * n_sites + col + 1] = acc[i][j][1];
        }
    }
}

struct FinishParams {
    double cell[9], cellinv[9];
    int pbc[3];
    double sqrt_alpha, r_cut, eps, self_term;
    int fourier_only;   // k-range shards other than part 0: reciprocal-space partial sum only
};

// Real-space term with the reference's 27-image minimum-image search
// (core.py:304-361), self term and the final combination (core.py:1594-1602,
// 1659-1661), fused with the split-k reduction.
__global__ void __launch_bounds__(256)
ewald_finish_kernel(const double *__restrict__ coords, long long n_sites, long long row0,
                    long long n_rows, FinishParams fp, const double *partials, int k_split,
                    double *out)
{
    const long long total = n_rows * n_sites;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long il = idx / n_sites, j = idx - il * n_sites, i = row0 + il;
        double f = 0.0;
        for (int s = 0; s < k_split; ++s) f += partials[(long long)s * total + idx];
        if (fp.fourier_only) {
            out[idx] = (f * 2) / fp.eps;
            continue;
        }
        const double dx = coords[3 * j] - coords[3 * i], dy = coords[3 * j + 1] - coords[3 * i + 1],
                     dz = coords[3 * j + 2] - coords[3 * i + 2];
        double fr[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            fr[k] = dx * fp.cellinv[k] + dy * fp.cellinv[3 + k] + dz * fp.cellinv[6 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (fp.pbc[k]) fr[k] -= rint(fr[k]);
        double best2 = 1e300;
#pragma unroll
        for (int ox = -1; ox <= 1; ++ox)
#pragma unroll
            for (int oy = -1; oy <= 1; ++oy)
#pragma unroll
                for (int oz = -1; oz <= 1; ++oz) {
                    const double tx = fr[0] + (fp.pbc[0] ? ox : 0), ty = fr[1] + (fp.pbc[1] ? oy : 0),
                                 tz = fr[2] + (fp.pbc[2] ? oz : 0);
                    const double x = tx * fp.cell[0] + ty * fp.cell[3] + tz * fp.cell[6];
                    const double y = tx * fp.cell[1] + ty * fp.cell[4] + tz * fp.cell[7];
                    const double z = tx * fp.cell[2] + ty * fp.cell[5] + tz * fp.cell[8];
                    const double r2 = x * x + y * y + z * z;
                    best2 = r2 < best2 ? r2 : best2;
                }
        const double r = sqrt(best2);
        double real = 0.0;
        if (r < fp.r_cut) real = (erfc(fp.sqrt_alpha * r) / 2) / (i == j ? 1.0 : r);
        out[idx] = real / fp.eps + (f * 2) / fp.eps + (i == j ? fp.self_term : 0.0);
    }
}

// P[i,j] = Pu[b_i][(cell_j - cell_i mod size)*nb + b_j].  HBM-write bound (8 N bytes per row): one CTA per
// (row i, x index of the target cell); for every y the sz*nb elements of the z column are contiguous in the
// output AND in the source row up to one wrap (source offset = target offset - z_i*nb, + sz*nb if negative),
// so an element costs a compare, two adds, a load from the L2-resident unit-cell rows and a coalesced store --
// no division per element (the first version spent its time in five 64-bit divisions per element).  The
// sy*sz*nb elements of one (row, x) are walked as ONE flat range (a z column of 300 elements would leave
// 212 of 512 thread slots idle when walked by itself); V = 2: 16-byte loads and stores when nb is even
// (every run, wrap point and row start is then an even offset).
template <int V>
__global__ void __launch_bounds__(256)
ewald_expand_kernel(const double *__restrict__ pu, int nb, int sx, int sy, int sz,
                    long long n_sites, long long row0, long long n_rows, double *__restrict__ out)
{
    using Vec = typename std::conditional<V == 2, double2, double>::type;
    const int xj = blockIdx.x;
    const int szn = sz * nb;
    const int szv = szn / V;                      // vector elements per z column
    const int per_x = sy * szv;
    const int q256 = 256 / szv, r256 = 256 - q256 * szv;
    for (long long il = blockIdx.y; il < n_rows; il += gridDim.y) {
        const int i = (int)(row0 + il);
        const int ci = i / nb, bi = i - ci * nb;
        const int zi = ci % sz, yi = (ci / sz) % sy, xi = ci / (sz * sy);
        int ddx = xj - xi;
        ddx += ddx < 0 ? sx : 0;
        const int shiftv = zi * nb / V;
        const Vec *src_x = reinterpret_cast<const Vec *>(pu + (long long)bi * n_sites + (long long)ddx * sy * szn);
        Vec *dst = reinterpret_cast<Vec *>(out + il * n_sites + (long long)xj * sy * szn);
        int yj = threadIdx.x / szv, e = threadIdx.x - yj * szv;
        for (int f = threadIdx.x; f < per_x; f += 256) {
            int ddy = yj - yi;
            ddy += ddy < 0 ? sy : 0;
            int se = e - shiftv;
            se += se < 0 ? szv : 0;
            dst[f] = __ldg(src_x + ddy * szv + se);
            e += r256;
            yj += q256;
            if (e >= szv) { e -= szv; ++yj; }
        }
    }
}

static void invert3(const double *m, double *inv) {
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    PYCD_REQUIRE(det != 0.0, "singular simulation cell");
    inv[0] = (e * i - f * h) / det; inv[1] = (c * h - b * i) / det; inv[2] = (b * f - c * e) / det;
    inv[3] = (f * g - d * i) / det; inv[4] = (a * i - c * g) / det; inv[5] = (c * d - a * f) / det;
    inv[6] = (d * h - e * g) / det; inv[7] = (b * g - a * h) / det; inv[8] = (a * e - b * d) / det;
}

// Half-space k list ordered in (n1,n2) columns with n3 ascending, padded to a
// multiple of kc with zero-weight entries.  Any half space gives the same sum
// (cos is even); the reference keeps the lexicographically negative triples
// (core.py:816-838), here the half space is n3>0 | (n3==0,n2>0) | (n3==n2==0,n1>0)
// so that runs along n3 are contiguous.
static int64_t build_k_list(const pycd_ewald_desc &d, int kc, std::vector<KEntry> &ent,
                            std::vector<double> &w) {
    const double alpha4 = 4 * d.alpha;
    const double coeff = (2 * M_PI) / d.volume;
    const double kc2 = d.k_cut * d.k_cut;
    PYCD_REQUIRE(d.k_max[0] < 32000 && d.k_max[1] < 32000 && d.k_max[2] < 32000, "k_max too large");
    ent.clear();
    w.clear();
    for (int n1 = -d.k_max[0]; n1 <= d.k_max[0]; ++n1)
        for (int n2 = -d.k_max[1]; n2 <= d.k_max[1]; ++n2) {
            int prev = INT32_MIN;
            for (int n3 = 0; n3 <= d.k_max[2]; ++n3) {
                if (n3 == 0 && !(n2 > 0 || (n2 == 0 && n1 > 0))) continue;
                double kv[3];
                for (int c = 0; c < 3; ++c)
                    kv[c] = n1 * d.recip[0 * 3 + c] + n2 * d.recip[1 * 3 + c] + n3 * d.recip[2 * 3 + c];
                const double k2 = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
                if (!(k2 < kc2)) continue;
                KEntry e;
                e.n1 = (short)n1; e.n2 = (short)n2; e.n3 = (short)n3;
                const bool cont = (prev == n3 - 1);
                e.flag = cont ? 1 : 0;
                ent.push_back(e);
                w.push_back(coeff * exp(-k2 / alpha4) / k2);  // core.py:869-875
                prev = n3;
            }
        }
    const int64_t k_eff = (int64_t)ent.size();
    while (ent.size() % (size_t)kc != 0) {
        KEntry e = {0, 0, 0, 0};
        ent.push_back(e);
        w.push_back(0.0);
    }
    return k_eff;
}

template <int BM, int BN, int TM, int TN, int KC, int CTAS>
static void launch_fourier(pycd_ctx *ctx, const double *theta, long long n, long long row0,
                           long long n_rows, const KEntry *kent, const double *kw, int n_chunks,
                           int k_split, double *out) {
    using T = EwaldTile<BM, BN, TM, TN, KC>;
    auto kern = ewald_fourier_kernel<BM, BN, TM, TN, KC, CTAS>;
    PYCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::smem_bytes));
    const int n_row_tiles = (int)((n_rows + BM - 1) / BM), n_col_tiles = (int)((n + BN - 1) / BN);
    const long long grid = (long long)k_split * n_row_tiles * n_col_tiles;
    PYCD_REQUIRE(grid < (1ll << 31), "grid too large");
    kern<<<(unsigned)grid, EW_THREADS, T::smem_bytes, ctx->stream>>>(
        theta, n, row0, n_rows, kent, kw, n_chunks, k_split, n_row_tiles, n_col_tiles, out);
    check_launch(ctx, "ewald_fourier_kernel");
}

template <int BM, int BN, int WM, int WN, int KC>
static void launch_fourier_dmma(pycd_ctx *ctx, const double *theta, long long n, long long row0,
                                long long n_rows, const KEntry *kent, const double *kw, int n_chunks,
                                int k_split, double *out) {
    auto kern = ewald_fourier_dmma_kernel<BM, BN, WM, WN, KC>;
    const size_t smem = sizeof(double) * (size_t)(2 * KC) * (BM + BN) + (sizeof(KEntry) + sizeof(double)) * KC;
    PYCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_row_tiles = (int)((n_rows + BM - 1) / BM), n_col_tiles = (int)((n + BN - 1) / BN);
    const long long grid = (long long)k_split * n_row_tiles * n_col_tiles;
    PYCD_REQUIRE(grid < (1ll << 31), "grid too large");
    kern<<<(unsigned)grid, EW_THREADS, smem, ctx->stream>>>(
        theta, n, row0, n_rows, kent, kw, n_chunks, k_split, n_row_tiles, n_col_tiles, out);
    check_launch(ctx, "ewald_fourier_dmma_kernel");
}

template <int BM, int BN, int WM, int WN, int KC, int CTAS>
static void launch_fourier_pipe(pycd_ctx *ctx, const double *theta, long long n, long long row0,
                                long long n_rows, const KEntry *kent, const double *kw, int n_chunks,
                                int k_split, double *out) {
    auto kern = ewald_fourier_pipe_kernel<BM, BN, WM, WN, KC, CTAS>;
    const size_t smem = 2 * (sizeof(double) * (size_t)(2 * KC) * (BM + BN) + (sizeof(KEntry) + sizeof(double)) * KC) + 16;
    PYCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_row_tiles = (int)((n_rows + BM - 1) / BM), n_col_tiles = (int)((n + BN - 1) / BN);
    const long long grid = (long long)k_split * n_row_tiles * n_col_tiles;
    PYCD_REQUIRE(grid < (1ll << 31), "grid too large");
    kern<<<(unsigned)grid, EW_THREADS, smem, ctx->stream>>>(
        theta, n, row0, n_rows, kent, kw, n_chunks, k_split, n_row_tiles, n_col_tiles, out);
    check_launch(ctx, "ewald_fourier_pipe_kernel");
}

// real-space + self terms + final combination of rows [0, n_rows) whose reciprocal-space sums are in
// `partials` (one partial): shared by the class-factorised unit-cell rows (ewald_cells.cu)
void ewald_finish_rows(pycd_ctx *ctx, const pycd_ewald_desc *desc, const double *coords_dev, long long n,
                       long long n_rows, const double *partials, double *out_dev) {
    FinishParams fp;
    for (int k = 0; k < 9; ++k) fp.cell[k] = desc->cell[k];
    invert3(desc->cell, fp.cellinv);
    for (int k = 0; k < 3; ++k) fp.pbc[k] = desc->pbc[k];
    fp.sqrt_alpha = sqrt(desc->alpha);
    fp.r_cut = desc->r_cut;
    fp.eps = desc->dielectric;
    fp.self_term = -sqrt(desc->alpha / M_PI) / desc->dielectric;  // core.py:1659
    fp.fourier_only = 0;
    const long long total = n_rows * n;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)ctx->n_sm * 16);
    ewald_finish_kernel<<<blocks, 256, 0, ctx->stream>>>(coords_dev, n, 0, n_rows, fp, partials, 1, out_dev);
    check_launch(ctx, "ewald_finish_kernel");
}

}  // namespace pycd

using namespace pycd;

extern "C" int pycd_ewald_rows(pycd_ctx *ctx, const pycd_ewald_desc *desc, int64_t row_begin,
                               int64_t row_end, double *out, pycd_ewald_stats *stats) {
    return guarded([&] {
        NvtxRange nvtx("pycd.ewald_rows");
        PYCD_REQUIRE(ctx && desc && out, "NULL argument");
        const long long n = desc->n_sites;
        PYCD_REQUIRE(n > 0 && row_begin >= 0 && row_end <= n && row_begin < row_end, "bad row range");
        PYCD_REQUIRE(desc->alpha > 0 && desc->volume > 0 && desc->dielectric > 0, "bad Ewald parameters");
        DeviceGuard g(ctx);
        const long long n_rows = row_end - row_begin;
        // rows the launch plan is derived from: the call's own, or -- row blocks of a sharded array -- those
        // of the whole array, so that every block sums the k vectors exactly like the one-GPU evaluation
        PYCD_REQUIRE(desc->plan_rows >= 0, "bad plan_rows");
        const long long plan_rows = desc->plan_rows > 0 ? std::max<long long>(desc->plan_rows, n_rows) : n_rows;
        // dense tile (64x128) for big row blocks, skinny tile (32x256) for the rows of one unit cell
        const bool wide = plan_rows > 64;
        // kernel variant (PYCD_EWALD_VARIANT, for A/B measurements): default = software-pipelined DMMA
        // kernel with 16-entry chunks (dense tile 64x128: 2 CTAs/SM; skinny tile 32x256: 1 CTA/SM with
        // 147 KB of double-buffered panels); "pipe16x1" / "pipe8" / "pipe8x1" = other chunk sizes / CTAs
        // per SM of it (dense tile only); "phased" = two-phase DMMA kernel; "dfma" = register-tiled DFMA
        const char *var_env = getenv("PYCD_EWALD_VARIANT");
        const std::string variant = var_env ? var_env : "";
        const bool use_dmma = variant != "dfma";
        const bool phased = variant == "phased";
        const bool pipe8 = variant == "pipe8" || variant == "pipe8x1";
        const bool pipe16 = use_dmma && !phased && !pipe8;
        const bool one_cta = variant == "pipe16x1" || variant == "pipe8x1";
        const int KC = (wide && !use_dmma) ? 32 : 16;   // the k list is padded to a multiple of KC

        std::vector<KEntry> ent;
        std::vector<double> w;
        const int64_t k_eff = build_k_list(*desc, KC, ent, w);
        const int n_chunks_all = (int)(ent.size() / KC);
        // k-range shard: part k_part of k_parts contiguous parts of the (padded) k list
        const int k_parts = desc->k_parts > 1 ? desc->k_parts : 1;
        PYCD_REQUIRE(desc->k_part >= 0 && desc->k_part < k_parts, "bad k_part");
        const int chunk_lo = (int)(((long long)n_chunks_all * desc->k_part) / k_parts);
        const int chunk_hi = (int)(((long long)n_chunks_all * (desc->k_part + 1)) / k_parts);
        const int n_chunks = chunk_hi - chunk_lo;

        InBuf<double> coords;
        coords.bind(desc->coords, (size_t)n * 3, ctx->stream);
        DevBuf<double> theta, kw;
        DevBuf<KEntry> kent;
        theta.alloc((size_t)n * 3);
        const double *B = desc->recip;
        ewald_theta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            coords.p, n, B[0], B[1], B[2], B[3], B[4], B[5], B[6], B[7], B[8], theta.p);
        check_launch(ctx, "ewald_theta_kernel");
        OutBuf<double> o;
        o.bind(out, (size_t)n_rows * n);

        int k_split = 1;
        if (n_chunks > 0) {
            kent.alloc((size_t)n_chunks * KC);
            kw.alloc((size_t)n_chunks * KC);
            PYCD_CUDA(cudaMemcpyAsync(kent.p, ent.data() + (size_t)chunk_lo * KC, (size_t)n_chunks * KC * sizeof(KEntry),
                                      cudaMemcpyHostToDevice, ctx->stream));
            PYCD_CUDA(cudaMemcpyAsync(kw.p, w.data() + (size_t)chunk_lo * KC, (size_t)n_chunks * KC * sizeof(double),
                                      cudaMemcpyHostToDevice, ctx->stream));
            const long long bm = wide ? (use_dmma ? 64 : 128) : 32, bn = wide ? 128 : 256;
            const long long tiles = ((plan_rows + bm - 1) / bm) * ((n + bn - 1) / bn);
            // split-k so that the grid fills whole waves of the SMs (1 CTA/SM): among the
            // candidates pick the best-filled last wave, preferring fewer splits on ties
            const long long ws_cap = std::max(1ll, (1ll << 30) / (plan_rows * n * 8));
            const long long ks_max = std::min({(long long)n_chunks, ws_cap, 32ll});
            const long long slots = (long long)ctx->n_sm * ((one_cta || ((pipe16 || pipe8) && !wide)) ? 1 : (use_dmma || !wide) ? 2 : 1);
            double best_fill = -1.0;
            for (long long ks = 1; ks <= ks_max; ++ks) {
                const long long grid = tiles * ks;
                const long long waves = (grid + slots - 1) / slots;
                const double fill = (double)grid / (double)(waves * slots);
                if (fill > best_fill + 0.02) { best_fill = fill; k_split = (int)ks; }
            }
        }
        DevBuf<double> ws;
        double *partials = o.dev();
        if (k_split > 1) {
            ws.alloc((size_t)k_split * n_rows * n);
            partials = ws.p;
        }
        KernelTimer tf(ctx, KC_EWALD_FOURIER);
        if (n_chunks > 0) {
            if (pipe16 && wide && one_cta)
                launch_fourier_pipe<64, 128, 32, 32, 16, 1>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                            n_chunks, k_split, partials);
            else if (pipe8 && wide && one_cta)
                launch_fourier_pipe<64, 128, 32, 32, 8, 1>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                           2 * n_chunks, k_split, partials);
            else if (pipe16 && wide)
                launch_fourier_pipe<64, 128, 32, 32, 16, 2>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                            n_chunks, k_split, partials);
            else if ((pipe16 || pipe8) && !wide)
                launch_fourier_pipe<32, 256, 32, 32, 16, 1>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                            n_chunks, k_split, partials);
            else if (pipe8 && wide)
                launch_fourier_pipe<64, 128, 32, 32, 8, 2>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                           2 * n_chunks, k_split, partials);
            else if (use_dmma && wide)
                launch_fourier_dmma<64, 128, 32, 32, 16>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                         n_chunks, k_split, partials);
            else if (use_dmma)
                launch_fourier_dmma<32, 256, 32, 32, 16>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                         n_chunks, k_split, partials);
            else if (wide)
                launch_fourier<128, 128, 8, 8, 32, 1>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                     n_chunks, k_split, partials);
            else
                launch_fourier<32, 256, 8, 4, 16, 2>(ctx, theta.p, n, row_begin, n_rows, kent.p, kw.p,
                                                    n_chunks, k_split, partials);
        } else {
            PYCD_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * n_rows * n, ctx->stream));
        }
        tf.stop(n_chunks > 0 ? 1 : 0);

        FinishParams fp;
        for (int k = 0; k < 9; ++k) fp.cell[k] = desc->cell[k];
        invert3(desc->cell, fp.cellinv);
        for (int k = 0; k < 3; ++k) fp.pbc[k] = desc->pbc[k];
        fp.sqrt_alpha = sqrt(desc->alpha);
        fp.r_cut = desc->r_cut;
        fp.eps = desc->dielectric;
        fp.self_term = -sqrt(desc->alpha / M_PI) / desc->dielectric;  // core.py:1659
        fp.fourier_only = desc->k_part > 0 && k_parts > 1;
        KernelTimer tr(ctx, KC_EWALD_FINISH);
        const long long total = n_rows * n;
        const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)ctx->n_sm * 16);
        ewald_finish_kernel<<<blocks, 256, 0, ctx->stream>>>(coords.p, n, row_begin, n_rows, fp, partials,
                                                             k_split, o.dev());
        check_launch(ctx, "ewald_finish_kernel");
        tr.stop(1);
        o.finish(ctx->stream);
        PYCD_CUDA(cudaStreamSynchronize(ctx->stream));
        tf.read();
        tr.read();
        if (stats) {
            stats->k_eff = k_eff;
            stats->rows = n_rows;
            stats->fourier_ms = ctx->last_ms[KC_EWALD_FOURIER];
            stats->finish_ms = ctx->last_ms[KC_EWALD_FINISH];
            stats->flops = 4.0 * (double)n_rows * (double)n * (double)k_eff + 40.0 * (double)n_rows * (double)n;
            stats->k_split = k_split;
        }
    });
}

extern "C" int pycd_ewald_expand(pycd_ctx *ctx, const double *p_unit, int32_t n_basis,
                                 const int32_t size[3], int64_t row_begin, int64_t row_end,
                                 double *out) {
    return guarded([&] {
        NvtxRange nvtx("pycd.ewald_expand");
        PYCD_REQUIRE(ctx && p_unit && size && out, "NULL argument");
        PYCD_REQUIRE(n_basis > 0 && size[0] > 0 && size[1] > 0 && size[2] > 0, "bad supercell");
        const long long n = (long long)n_basis * size[0] * size[1] * size[2];
        PYCD_REQUIRE(n < (1ll << 31), "too many sites");
        PYCD_REQUIRE(row_begin >= 0 && row_end <= n && row_begin < row_end, "bad row range");
        DeviceGuard g(ctx);
        const long long n_rows = row_end - row_begin;
        InBuf<double> pu;
        pu.bind(p_unit, (size_t)n_basis * n, ctx->stream);
        OutBuf<double> o;
        o.bind(out, (size_t)n_rows * n);
        KernelTimer t(ctx, KC_EWALD_EXPAND);
        const dim3 grid((unsigned)size[0], (unsigned)std::min<long long>(n_rows, 65535));
        const bool vec2 = n_basis % 2 == 0 && ((uintptr_t)pu.p % 16) == 0 && ((uintptr_t)o.dev() % 16) == 0;
        if (vec2)
            ewald_expand_kernel<2><<<grid, 256, 0, ctx->stream>>>(pu.p, n_basis, size[0], size[1], size[2], n,
                                                                  row_begin, n_rows, o.dev());
        else
            ewald_expand_kernel<1><<<grid, 256, 0, ctx->stream>>>(pu.p, n_basis, size[0], size[1], size[2], n,
                                                                  row_begin, n_rows, o.dev());
        check_launch(ctx, "ewald_expand_kernel");
        t.stop(1);
        o.finish(ctx->stream);
        PYCD_CUDA(cudaStreamSynchronize(ctx->stream));
        t.read();
    });
}
