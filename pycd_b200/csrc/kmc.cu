// kmc.cu -- batched lattice-KMC step kernel on sm_100a: rate evaluation
// (PyCD/core.py:1946-2050), event selection + time advance + recording
// (core.py:2796-2861) for many independent trajectories, one CTA per trajectory.
//
// Electrostatics use the O(C) gather identity instead of the reference's O(N) dot
// (SURVEY A.4): with q = q_lat + sum_c q_c e_{s_c},
//   q.(P[b,:]-P[a,:]) = (V_lat[b]-V_lat[a]) + sum_c' q_c (P[b,s_c'] - P[a,s_c']),  V_lat = P.q_lat
// refresh_interval = 1 re-gathers every process every step (the stateless
// formulation the roofline figure B_step counts); R > 1 patches the cached sums
// after each hop (4 gathers per untouched process) and re-gathers every R steps.
#include "common.cuh"

#include <algorithm>
#include <climits>
#include <cmath>

namespace pycd {

struct SysDev {
    const double *P;
    long long n_sites;
    const int *site_centre;
    const int *site_class;
    int nn;
    const int *neigh;
    const double *hopvec;
    const double *lam;
    const double *vab;
    const double *e_rel;
    const double *v_lat;
    double qc, kT, vn;
    double field[3];
    int field_active;
};

struct EnsDev {
    int C, n_proc;
    long long n_traj;
    unsigned long long traj_id0;
    int *occ;              // [n_traj][C]
    double *t;             // [n_traj]
    long long *start_idx;  // [n_traj]
    int *done;             // [n_traj]
    double *disp;          // [n_traj][3C] hops since the last recorded row
    double *row;           // [n_traj][3C] last recorded row
    long long *n_steps;    // [n_traj]
    long long *near_tie;   // [n_traj]
    long long *clamped;    // [n_traj]
    double *drift;         // [n_traj][3C]
    double *rates;         // [n_traj][n_proc]
    const double *kT_traj;
    const double *field_traj;
    double *unwrapped;     // [n_traj][n_path][3C] or NULL
    double dt_grid;
    long long n_path;
    long long step_limit;
    int stop_at_grid_end;
    int rng_mode;
    unsigned long long seed;
    int refresh_interval;
};

struct AdvanceArgs {
    long long max_steps;
    const double *draws;   // [n_traj][2*max_steps]
    int *events_out;       // [n_traj][max_steps]
    double *times_out;     // [n_traj][max_steps]
    long long *steps_done; // [n_traj]
};

// Philox4x32-10; identical to oracle/pycd_oracle.c (tests/test_philox.py)
__device__ __forceinline__ void philox_uniforms(unsigned long long seed, unsigned long long traj,
                                                unsigned long long step, double &u1, double &u2)
{
    unsigned int c0 = (unsigned int)step, c1 = (unsigned int)(step >> 32);
    unsigned int c2 = (unsigned int)traj, c3 = (unsigned int)(traj >> 32);
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const unsigned int h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const unsigned long long a = ((unsigned long long)(c0 >> 5) << 26) + (c1 >> 6);
    const unsigned long long b = ((unsigned long long)(c2 >> 5) << 26) + (c3 >> 6);
    u1 = (double)a * (1.0 / 9007199254740992.0);
    u2 = (double)(b + 1) * (1.0 / 9007199254740992.0);
}

constexpr double TIE_TOL = 1e-12;  // >> n_proc*eps: scan-order differences cannot cross it

// Per-process cached quantities live in shared memory for the whole launch.
struct ProcSmem {
    int *a, *b;                 // old / new site
    double *t01, *t02, *shift;  // core.py:2004-2014, 2023-2025
    double *lam, *vab, *fs;     // lambda, V_AB, 0.5*E.hop
    double *k, *cum;
};

__device__ __forceinline__ void gather_process(const SysDev &S, const int *s_occ, int C, int p,
                                               const double *fld, int field_active, ProcSmem &M)
{
    const int c = p / S.nn, slot = p - c * S.nn;
    const int a = s_occ[c];
    const int e = S.site_centre[a];
    const int b = S.neigh[(long long)e * S.nn + slot];
    const int cls = S.site_class[a];
    const double *Pa = S.P + (long long)a * S.n_sites, *Pb = S.P + (long long)b * S.n_sites;
    // term01 in the oracle's order: start from the lattice part, add carriers in order
    double t01 = __dsub_rn(S.v_lat[b], S.v_lat[a]);
#pragma unroll 4
    for (int c2 = 0; c2 < C; ++c2) {
        const int sc = s_occ[c2];
        t01 = __dadd_rn(t01, __dmul_rn(S.qc, __dsub_rn(__ldg(Pb + sc), __ldg(Pa + sc))));
    }
    M.a[p] = a;
    M.b[p] = b;
    M.t01[p] = t01;
    M.t02[p] = __dmul_rn(S.qc, __dsub_rn(__ldg(Pa + a), __ldg(Pa + b)));
    M.shift[p] = __dsub_rn(S.e_rel[b], S.e_rel[a]);
    M.lam[p] = S.lam[cls * S.nn + slot];
    M.vab[p] = S.vab[cls * S.nn + slot];
    double fs = 0.0;
    if (field_active) {
        const double *hv = S.hopvec + ((long long)e * S.nn + slot) * 3;
        fs = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(fld[0], hv[0]), __dmul_rn(fld[1], hv[1])),
                                      __dmul_rn(fld[2], hv[2])));
    }
    M.fs[p] = fs;
}

template <int BS>
__global__ void __launch_bounds__(BS)
kmc_step_kernel(SysDev S, EnsDev E, AdvanceArgs A)
{
    const int traj = blockIdx.x;
    const int tid = threadIdx.x;
    const int C = E.C, n_proc = E.n_proc, nn = S.nn;
    constexpr int NW = BS / 32;
    const int lane = tid & 31, wid = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sd = reinterpret_cast<double *>(smem_raw);
    ProcSmem M;
    M.t01 = sd; sd += n_proc;
    M.t02 = sd; sd += n_proc;
    M.shift = sd; sd += n_proc;
    M.lam = sd; sd += n_proc;
    M.vab = sd; sd += n_proc;
    M.fs = sd; sd += n_proc;
    M.k = sd; sd += n_proc;
    M.cum = sd; sd += n_proc;
    double *s_disp = sd; sd += 3 * C;
    double *s_row = sd; sd += 3 * C;
    double *s_drift = sd; sd += 3 * C;
    double *s_wsum = sd; sd += 32;
    long long *s_ll = reinterpret_cast<long long *>(sd); sd += 2;  // rows [s_ll[0], s_ll[1]) pending
    int *si = reinterpret_cast<int *>(sd);
    M.a = si; si += n_proc;
    M.b = si; si += n_proc;
    int *s_occ = si; si += C;
    int *s_flag = si; si += 2;      // [0] = selected process, [1] = trajectory finished

    if (E.done[traj]) {
        if (tid == 0 && A.steps_done) A.steps_done[traj] = 0;
        return;
    }

    double kT = E.kT_traj ? E.kT_traj[traj] : S.kT;
    double fld[3] = {S.field[0], S.field[1], S.field[2]};
    int field_active = S.field_active;
    if (E.field_traj) {
        fld[0] = E.field_traj[3 * traj];
        fld[1] = E.field_traj[3 * traj + 1];
        fld[2] = E.field_traj[3 * traj + 2];
        field_active = (fld[0] != 0.0 || fld[1] != 0.0 || fld[2] != 0.0);
    }

    for (int c = tid; c < C; c += BS) s_occ[c] = E.occ[(long long)traj * C + c];
    for (int d = tid; d < 3 * C; d += BS) {
        s_disp[d] = E.disp[(long long)traj * 3 * C + d];
        s_row[d] = E.row[(long long)traj * 3 * C + d];
        s_drift[d] = E.drift[(long long)traj * 3 * C + d];
    }
    // thread-0 scalars
    double t = E.t[traj];
    long long start = E.start_idx[traj];
    long long steps_total = E.n_steps[traj];
    long long n_tie = 0, n_clamp = 0;
    long long step_local = 0;
    const unsigned long long traj_gid = E.traj_id0 + (unsigned long long)traj;
    const int R = E.refresh_interval;
    if (tid == 0) { s_flag[1] = 0; s_ll[0] = 0; s_ll[1] = 0; }
    __syncthreads();

    while (true) {
        // ---- pending recording of the previous step (rows [rec_start, rec_end)) ----
        {
            const long long r0 = s_ll[0], r1 = s_ll[1];
            if (r1 > r0) {
                // unwrapped[start:end] = unwrapped[start-1] + displacement, core.py:2852-2854
                for (int d = tid; d < 3 * C; d += BS) {
                    const double v = s_row[d] + s_disp[d];
                    s_row[d] = v;
                    s_disp[d] = 0.0;
                    if (E.unwrapped) {
                        double *dst = E.unwrapped + ((long long)traj * E.n_path + r0) * 3 * C + d;
                        for (long long r = r0; r < r1; ++r, dst += 3 * C) *dst = v;
                    }
                }
            }
        }
        if (s_flag[1] || step_local >= A.max_steps) break;
        __syncthreads();  // everyone has read s_ll / s_flag before thread 0 rewrites them

        // ---- rates ----
        const bool full = (R <= 1) || ((steps_total + step_local) % R == 0);
        double ksum = 0.0;
        for (int p = tid; p < n_proc; p += BS) {
            if (full) gather_process(S, s_occ, C, p, fld, field_active, M);
            const double ew = __dmul_rn(__dmul_rn(2.0, S.qc), __dadd_rn(M.t01[p], M.t02[p]));  // core.py:2016
            const double g0 = __dadd_rn(ew, M.shift[p]);
            const double lam = M.lam[p];
            const double lg = __dadd_rn(lam, g0);
            const double gs = __dsub_rn(__dsub_rn(__ddiv_rn(__dmul_rn(lg, lg), __dmul_rn(4.0, lam)), M.vab[p]),
                                        M.fs[p]);                                                // core.py:2045
            const double kp = __dmul_rn(S.vn, pow(2.718281828459045, __ddiv_rn(-gs, kT)));        // core.py:2047
            M.k[p] = kp;
            ksum += kp;
        }
        // block sum (fixed tree order)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ksum += __shfl_xor_sync(0xffffffffu, ksum, o);
        if (lane == 0) s_wsum[wid] = ksum;
        __syncthreads();
        double ktot = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) ktot += s_wsum[w];
        __syncthreads();  // s_wsum reused by the scan

        // ---- normalised inclusive scan, core.py:2797 ----
        double carry = 0.0;
        for (int base = 0; base < n_proc; base += BS) {
            const int p = base + tid;
            double x = (p < n_proc) ? __ddiv_rn(M.k[p], ktot) : 0.0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wsum[wid] = x;
            __syncthreads();
            double pre = carry;
            for (int w = 0; w < wid; ++w) pre += s_wsum[w];
            if (p < n_proc) M.cum[p] = pre + x;
            double tot = carry;
            for (int w = 0; w < NW; ++w) tot += s_wsum[w];
            carry = tot;
            __syncthreads();
        }

        // ---- uniform draws ----
        double u1, u2;
        if (E.rng_mode == PYCD_RNG_REPLAY) {
            const double *dr = A.draws + ((long long)traj * A.max_steps + step_local) * 2;
            u1 = dr[0];
            u2 = dr[1];
        } else {
            philox_uniforms(E.seed, traj_gid, (unsigned long long)(steps_total + step_local), u1, u2);
        }

        // ---- first index with cum > u1, core.py:2800 ----
        if (tid == 0) s_flag[0] = INT_MAX;
        __syncthreads();
        for (int p = tid; p < n_proc; p += BS)
            if (M.cum[p] > u1 && (p == 0 || !(M.cum[p - 1] > u1))) atomicMin(&s_flag[0], p);
        __syncthreads();
        int sel = s_flag[0];
        bool tie = (sel == INT_MAX);
        if (!tie) {
            const double hi = M.cum[sel], lo = sel > 0 ? M.cum[sel - 1] : 0.0;
            tie = (hi - u1 < TIE_TOL) || (sel > 0 && u1 - lo < TIE_TOL);
        }
        if (tie) {  // block-uniform branch: redo the selection in the reference's sequential order
            __syncthreads();
            if (tid == 0) {
                double kseq = 0.0;
                for (int p = 0; p < n_proc; ++p) kseq += M.k[p];
                double cum = 0.0;
                int s2 = -1;
                for (int p = 0; p < n_proc; ++p) {
                    cum += M.k[p] / kseq;
                    if (cum > u1) { s2 = p; break; }
                }
                if (s2 < 0) { s2 = n_proc - 1; ++n_clamp; }  // reference raises IndexError here
                ++n_tie;
                s_flag[0] = s2;
            }
            __syncthreads();
            sel = s_flag[0];
        }

        const int cs = sel / nn, slot = sel - cs * nn;
        const int a_old = M.a[sel], b_new = M.b[sel];
        const bool next_full = (R <= 1) || ((steps_total + step_local + 1) % R == 0);

        // ---- issue the cache patches first (long-latency gathers) ----
        double patch[4];
        int npatch = 0;
        if (!next_full) {
            for (int p = tid; p < n_proc && npatch < 4; p += BS) {
                if (p / nn == cs) { patch[npatch++] = 0.0; continue; }
                const long long ra = (long long)M.a[p] * S.n_sites, rb = (long long)M.b[p] * S.n_sites;
                const double nb_ = __ldg(S.P + rb + b_new), na_ = __ldg(S.P + ra + b_new);
                const double ob_ = __ldg(S.P + rb + a_old), oa_ = __ldg(S.P + ra + a_old);
                patch[npatch++] = S.qc * (nb_ - na_) - S.qc * (ob_ - oa_);
            }
        }

        // ---- thread 0: time advance + bookkeeping, core.py:2802-2830, 2844-2861 ----
        if (tid == 0) {
            t -= log(u2) / ktot;
            const long long end = (long long)(t / E.dt_grid);
            const int e = S.site_centre[a_old];
            const double *hv = S.hopvec + ((long long)e * nn + slot) * 3;
            const double kp = M.k[sel];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                s_disp[3 * cs + d] += hv[d];
                if (field_active) s_drift[3 * cs + d] += hv[d] * kp;
            }
            if (A.events_out) A.events_out[(long long)traj * A.max_steps + step_local] = sel;
            if (A.times_out) A.times_out[(long long)traj * A.max_steps + step_local] = t;
            long long r0 = 0, r1 = 0;
            int fin = 0;
            if (end >= start + 1) {  // core.py:2848-2861
                const long long e2 = end >= E.n_path ? E.n_path : end;
                if (start < E.n_path) { r0 = start; r1 = e2; }
                start = e2;
            }
            if (E.stop_at_grid_end && end >= E.n_path) fin = 1;
            if (E.step_limit > 0 && steps_total + step_local + 1 >= E.step_limit) fin = 1;
            s_ll[0] = r0;
            s_ll[1] = r1;
            s_flag[1] = fin;
        }
        if (tid == 0) s_occ[cs] = b_new;
        __syncthreads();  // s_occ, s_ll, s_flag visible

        // ---- bring the cached sums up to date for the next step ----
        if (!next_full) {
            int ip = 0;
            for (int p = tid; p < n_proc; p += BS, ++ip) {
                if (p / nn == cs) {
                    gather_process(S, s_occ, C, p, fld, field_active, M);
                } else if (ip < 4) {
                    M.t01[p] += patch[ip];
                } else {
                    const long long ra = (long long)M.a[p] * S.n_sites, rb = (long long)M.b[p] * S.n_sites;
                    M.t01[p] += S.qc * (__ldg(S.P + rb + b_new) - __ldg(S.P + ra + b_new)) -
                                S.qc * (__ldg(S.P + rb + a_old) - __ldg(S.P + ra + a_old));
                }
            }
        }
        ++step_local;
        __syncthreads();
    }

    // ---- write the state back ----
    __syncthreads();
    for (int c = tid; c < C; c += BS) E.occ[(long long)traj * C + c] = s_occ[c];
    for (int d = tid; d < 3 * C; d += BS) {
        E.disp[(long long)traj * 3 * C + d] = s_disp[d];
        E.row[(long long)traj * 3 * C + d] = s_row[d];
        E.drift[(long long)traj * 3 * C + d] = s_drift[d];
    }
    if (step_local > 0)
        for (int p = tid; p < n_proc; p += BS) E.rates[(long long)traj * n_proc + p] = M.k[p];
    if (tid == 0) {
        E.t[traj] = t;
        E.start_idx[traj] = start;
        E.n_steps[traj] = steps_total + step_local;
        E.near_tie[traj] += n_tie;
        E.clamped[traj] += n_clamp;
        if (s_flag[1]) E.done[traj] = 1;
        if (A.steps_done) A.steps_done[traj] = step_local;
    }
}

// V_lat = P . q_lat, one warp per row, double-double accumulation so that the
// differences V_lat[b]-V_lat[a] keep ~1e-16 Ha accuracy at N = 30 000.
__global__ void __launch_bounds__(256)
vlat_kernel(const double *__restrict__ P, const double *__restrict__ q, long long n,
            double *__restrict__ v)
{
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double *pr = P + row * n;
    double hi = 0.0, lo = 0.0;
    for (long long j = lane; j < n; j += 32) {
        const double x = pr[j] * q[j];
        const double e = fma(pr[j], q[j], -x);  // exact product error
        const double s = hi + x;                // two-sum
        const double bb = s - hi;
        lo += ((hi - (s - bb)) + (x - bb)) + e;
        hi = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double h2 = __shfl_xor_sync(0xffffffffu, hi, o), l2 = __shfl_xor_sync(0xffffffffu, lo, o);
        const double s = hi + h2;
        const double bb = s - hi;
        lo += ((hi - (s - bb)) + (h2 - bb)) + l2;
        hi = s;
    }
    if (lane == 0) v[row] = hi + lo;
}

static size_t kmc_smem_bytes(int n_proc, int C) {
    const size_t doubles = (size_t)8 * n_proc + (size_t)9 * C + 32 + 2;
    const size_t ints = (size_t)2 * n_proc + C + 2;
    return doubles * 8 + ((ints + 1) / 2) * 8;
}

}  // namespace pycd

using namespace pycd;

struct pycd_kmc_system {
    pycd_ctx *ctx = nullptr;
    SysDev dev{};
    long long n_centres = 0;
    int n_class = 0;
    InBuf<double> P;
    InBuf<int> site_centre, site_class, neigh;
    InBuf<double> hopvec, lam, vab, e_rel;
    DevBuf<double> v_lat;
};

struct pycd_kmc_ensemble {
    pycd_kmc_system *sys = nullptr;
    EnsDev dev{};
    DevBuf<int> occ, done;
    DevBuf<double> t, disp, row, drift, rates, unwrapped, kT_traj, field_traj;
    DevBuf<long long> start_idx, n_steps, near_tie, clamped;
};

extern "C" int pycd_kmc_system_create(pycd_ctx *ctx, const pycd_kmc_system_desc *d,
                                      pycd_kmc_system **out) {
    return guarded([&] {
        PYCD_REQUIRE(ctx && d && out, "NULL argument");
        *out = nullptr;
        PYCD_REQUIRE(d->n_sites > 0 && d->n_sites < (1ll << 31), "bad n_sites");
        PYCD_REQUIRE(d->n_centres > 0 && d->nn > 0 && d->n_class > 0, "bad table sizes");
        PYCD_REQUIRE(d->P && d->site_centre && d->site_class && d->neigh && d->hopvec && d->lam &&
                         d->vab && d->e_rel && d->q_lat, "NULL table");
        PYCD_REQUIRE(d->kT > 0 && d->vn > 0, "bad kT / vn");
        DeviceGuard g(ctx);
        auto sys = new pycd_kmc_system();
        try {
            sys->ctx = ctx;
            const size_t n = (size_t)d->n_sites;
            cudaStream_t s = ctx->stream;
            sys->P.bind(d->P, n * n, s);
            sys->site_centre.bind(d->site_centre, n, s);
            sys->site_class.bind(d->site_class, n, s);
            sys->neigh.bind(d->neigh, (size_t)d->n_centres * d->nn, s);
            sys->hopvec.bind(d->hopvec, (size_t)d->n_centres * d->nn * 3, s);
            sys->lam.bind(d->lam, (size_t)d->n_class * d->nn, s);
            sys->vab.bind(d->vab, (size_t)d->n_class * d->nn, s);
            sys->e_rel.bind(d->e_rel, n, s);
            InBuf<double> q;
            q.bind(d->q_lat, n, s);
            sys->v_lat.alloc(n);
            KernelTimer tv(ctx, KC_VLAT);
            vlat_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(sys->P.p, q.p, (long long)n,
                                                                         sys->v_lat.p);
            check_launch(ctx, "vlat_kernel");
            tv.stop(1);
            PYCD_CUDA(cudaStreamSynchronize(s));
            tv.read();
            SysDev &v = sys->dev;
            v.P = sys->P.p; v.n_sites = d->n_sites;
            v.site_centre = sys->site_centre.p; v.site_class = sys->site_class.p;
            v.nn = d->nn; v.neigh = sys->neigh.p; v.hopvec = sys->hopvec.p;
            v.lam = sys->lam.p; v.vab = sys->vab.p; v.e_rel = sys->e_rel.p; v.v_lat = sys->v_lat.p;
            v.qc = d->q_carrier; v.kT = d->kT; v.vn = d->vn;
            for (int k = 0; k < 3; ++k) v.field[k] = d->field[k];
            v.field_active = d->field_active;
            sys->n_centres = d->n_centres;
            sys->n_class = d->n_class;
        } catch (...) {
            delete sys;
            throw;
        }
        *out = sys;
    });
}

extern "C" int pycd_kmc_system_destroy(pycd_kmc_system *sys) {
    return guarded([&] {
        if (!sys) return;
        DeviceGuard g(sys->ctx);
        delete sys;
    });
}

extern "C" int pycd_kmc_system_vlat(pycd_kmc_system *sys, double *v_lat) {
    return guarded([&] {
        PYCD_REQUIRE(sys && v_lat, "NULL argument");
        DeviceGuard g(sys->ctx);
        PYCD_CUDA(cudaMemcpy(v_lat, sys->v_lat.p, sizeof(double) * sys->dev.n_sites,
                             is_device_pointer(v_lat) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    });
}

extern "C" int pycd_kmc_ensemble_create(pycd_kmc_system *sys, const pycd_kmc_ensemble_desc *d,
                                        pycd_kmc_ensemble **out) {
    return guarded([&] {
        PYCD_REQUIRE(sys && d && out, "NULL argument");
        *out = nullptr;
        PYCD_REQUIRE(d->n_traj > 0 && d->n_traj < (1ll << 31), "bad n_traj");
        PYCD_REQUIRE(d->n_carriers > 0 && d->occupancy0, "bad carriers");
        PYCD_REQUIRE(d->dt_grid > 0 && d->n_path >= 1, "bad time grid");
        PYCD_REQUIRE(d->refresh_interval >= 1, "refresh_interval must be >= 1");
        PYCD_REQUIRE(d->rng_mode == PYCD_RNG_REPLAY || d->rng_mode == PYCD_RNG_PHILOX, "bad rng_mode");
        PYCD_REQUIRE(d->stop_at_grid_end || d->step_limit > 0, "trajectory would never end");
        pycd_ctx *ctx = sys->ctx;
        DeviceGuard g(ctx);
        const long long nt = d->n_traj;
        const int C = d->n_carriers;
        const long long n_proc_ll = (long long)C * sys->dev.nn;
        PYCD_REQUIRE(n_proc_ll <= 2560, "more than 2560 processes per trajectory is not supported");
        // validate initial sites on the host copy (they index device tables)
        std::vector<int> occ_h((size_t)nt * C);
        PYCD_CUDA(cudaMemcpy(occ_h.data(), d->occupancy0, sizeof(int) * nt * C,
                             is_device_pointer(d->occupancy0) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
        std::vector<int> sc((size_t)sys->dev.n_sites);
        PYCD_CUDA(cudaMemcpy(sc.data(), sys->dev.site_centre, sizeof(int) * sys->dev.n_sites, cudaMemcpyDeviceToHost));
        for (int v : occ_h)
            PYCD_REQUIRE(v >= 0 && v < sys->dev.n_sites && sc[v] >= 0, "initial occupancy is not a carrier site");
        auto ens = new pycd_kmc_ensemble();
        try {
            ens->sys = sys;
            cudaStream_t s = ctx->stream;
            ens->occ.alloc((size_t)nt * C);
            PYCD_CUDA(cudaMemcpyAsync(ens->occ.p, occ_h.data(), sizeof(int) * nt * C, cudaMemcpyHostToDevice, s));
            ens->done.alloc(nt); ens->done.zero(s);
            ens->t.alloc(nt); ens->t.zero(s);
            ens->disp.alloc((size_t)nt * 3 * C); ens->disp.zero(s);
            ens->row.alloc((size_t)nt * 3 * C); ens->row.zero(s);
            ens->drift.alloc((size_t)nt * 3 * C); ens->drift.zero(s);
            ens->rates.alloc((size_t)nt * n_proc_ll); ens->rates.zero(s);
            ens->n_steps.alloc(nt); ens->n_steps.zero(s);
            ens->near_tie.alloc(nt); ens->near_tie.zero(s);
            ens->clamped.alloc(nt); ens->clamped.zero(s);
            ens->start_idx.alloc(nt);
            std::vector<long long> ones((size_t)nt, 1);  // start_path_index = 1, core.py:2781
            PYCD_CUDA(cudaMemcpyAsync(ens->start_idx.p, ones.data(), sizeof(long long) * nt, cudaMemcpyHostToDevice, s));
            if (d->record_unwrapped) {
                ens->unwrapped.alloc((size_t)nt * d->n_path * 3 * C);
                ens->unwrapped.zero(s);
            }
            if (d->kT_traj) {
                ens->kT_traj.alloc(nt);
                PYCD_CUDA(cudaMemcpyAsync(ens->kT_traj.p, d->kT_traj, sizeof(double) * nt, cudaMemcpyDefault, s));
            }
            if (d->field_traj) {
                ens->field_traj.alloc((size_t)nt * 3);
                PYCD_CUDA(cudaMemcpyAsync(ens->field_traj.p, d->field_traj, sizeof(double) * nt * 3, cudaMemcpyDefault, s));
            }
            PYCD_CUDA(cudaStreamSynchronize(s));
            EnsDev &e = ens->dev;
            e.C = C; e.n_proc = (int)n_proc_ll; e.n_traj = nt; e.traj_id0 = d->traj_id0;
            e.occ = ens->occ.p; e.t = ens->t.p; e.start_idx = ens->start_idx.p; e.done = ens->done.p;
            e.disp = ens->disp.p; e.row = ens->row.p; e.n_steps = ens->n_steps.p;
            e.near_tie = ens->near_tie.p; e.clamped = ens->clamped.p; e.drift = ens->drift.p;
            e.rates = ens->rates.p; e.kT_traj = ens->kT_traj.p; e.field_traj = ens->field_traj.p;
            e.unwrapped = ens->unwrapped.p; e.dt_grid = d->dt_grid; e.n_path = d->n_path;
            e.step_limit = d->step_limit; e.stop_at_grid_end = d->stop_at_grid_end;
            e.rng_mode = d->rng_mode; e.seed = d->seed; e.refresh_interval = d->refresh_interval;
        } catch (...) {
            delete ens;
            throw;
        }
        *out = ens;
    });
}

extern "C" int pycd_kmc_ensemble_destroy(pycd_kmc_ensemble *ens) {
    return guarded([&] {
        if (!ens) return;
        DeviceGuard g(ens->sys->ctx);
        delete ens;
    });
}

template <int BS>
static void launch_step(pycd_ctx *ctx, const SysDev &S, const EnsDev &E, const AdvanceArgs &A, size_t smem) {
    auto kern = kmc_step_kernel<BS>;
    if (smem > 48 * 1024)
        PYCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)E.n_traj, BS, smem, ctx->stream>>>(S, E, A);
    check_launch(ctx, "kmc_step_kernel");
}

extern "C" int pycd_kmc_advance(pycd_kmc_ensemble *ens, int64_t max_steps, const double *draws,
                                int32_t *events_out, double *times_out, int64_t *steps_done,
                                int64_t *n_active) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        PYCD_REQUIRE(max_steps > 0, "max_steps must be positive");
        pycd_ctx *ctx = ens->sys->ctx;
        DeviceGuard g(ctx);
        EnsDev &E = ens->dev;
        PYCD_REQUIRE(E.refresh_interval <= 1 || max_steps % E.refresh_interval == 0,
                     "max_steps must be a multiple of refresh_interval");
        PYCD_REQUIRE(E.rng_mode != PYCD_RNG_REPLAY || draws, "REPLAY mode needs draws");
        const size_t nt = (size_t)E.n_traj;
        cudaStream_t s = ctx->stream;
        InBuf<double> dr;
        if (E.rng_mode == PYCD_RNG_REPLAY) dr.bind(draws, nt * 2 * (size_t)max_steps, s);
        OutBuf<int> ev;
        ev.bind(events_out, nt * (size_t)max_steps);
        OutBuf<double> tm;
        tm.bind(times_out, nt * (size_t)max_steps);
        OutBuf<long long> sdn;
        sdn.bind(reinterpret_cast<long long *>(steps_done), nt);
        if (ev.own.p) PYCD_CUDA(cudaMemsetAsync(ev.own.p, 0xff, sizeof(int) * nt * max_steps, s));
        if (tm.own.p) PYCD_CUDA(cudaMemsetAsync(tm.own.p, 0, sizeof(double) * nt * max_steps, s));
        AdvanceArgs A;
        A.max_steps = max_steps;
        A.draws = dr.p;
        A.events_out = ev.dev();
        A.times_out = tm.dev();
        A.steps_done = sdn.dev();
        const size_t smem = kmc_smem_bytes(E.n_proc, E.C);
        PYCD_REQUIRE(smem <= 200 * 1024, "trajectory state does not fit in shared memory");
        KernelTimer tk(ctx, KC_KMC_STEP);
        if (E.n_proc <= 32) launch_step<32>(ctx, ens->sys->dev, E, A, smem);
        else if (E.n_proc <= 64) launch_step<64>(ctx, ens->sys->dev, E, A, smem);
        else if (E.n_proc <= 128) launch_step<128>(ctx, ens->sys->dev, E, A, smem);
        else launch_step<256>(ctx, ens->sys->dev, E, A, smem);
        tk.stop(1);
        ev.finish(s);
        tm.finish(s);
        sdn.finish(s);
        std::vector<int> done_h(nt);
        PYCD_CUDA(cudaMemcpyAsync(done_h.data(), ens->done.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, s));
        PYCD_CUDA(cudaStreamSynchronize(s));
        tk.read();
        if (n_active) {
            int64_t act = 0;
            for (int v : done_h) act += (v == 0);
            *n_active = act;
        }
    });
}

extern "C" int pycd_kmc_read(pycd_kmc_ensemble *ens, double *unwrapped, int64_t *n_steps,
                             double *sim_time, int32_t *occupancy, double *drift, int64_t *near_tie,
                             int64_t *clamped, double *rates) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        pycd_ctx *ctx = ens->sys->ctx;
        DeviceGuard g(ctx);
        const EnsDev &E = ens->dev;
        const size_t nt = (size_t)E.n_traj;
        cudaStream_t s = ctx->stream;
        auto pull = [&](void *dst, const void *src, size_t bytes) {
            if (dst) PYCD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
        };
        if (unwrapped) {
            PYCD_REQUIRE(ens->unwrapped.p, "ensemble was created without record_unwrapped");
            pull(unwrapped, ens->unwrapped.p, sizeof(double) * nt * E.n_path * 3 * E.C);
        }
        pull(n_steps, ens->n_steps.p, sizeof(long long) * nt);
        pull(sim_time, ens->t.p, sizeof(double) * nt);
        pull(occupancy, ens->occ.p, sizeof(int) * nt * E.C);
        pull(drift, ens->drift.p, sizeof(double) * nt * 3 * E.C);
        pull(near_tie, ens->near_tie.p, sizeof(long long) * nt);
        pull(clamped, ens->clamped.p, sizeof(long long) * nt);
        pull(rates, ens->rates.p, sizeof(double) * nt * E.n_proc);
        PYCD_CUDA(cudaStreamSynchronize(s));
    });
}

extern "C" int pycd_kmc_unwrapped_device(pycd_kmc_ensemble *ens, const double **dev_ptr) {
    return guarded([&] {
        PYCD_REQUIRE(ens && dev_ptr, "NULL argument");
        PYCD_REQUIRE(ens->unwrapped.p, "ensemble was created without record_unwrapped");
        *dev_ptr = ens->unwrapped.p;
    });
}
