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
#include <cstring>
#include "kmc_types.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

namespace pycd {


// A lattice site as the kernel carries it: index + (compact layout) packed cell/basis.
struct Site {
    int idx;
    unsigned pack;
};

template <bool COMPACT>
__device__ __forceinline__ Site make_site(const SysDev &S, int idx)
{
    Site s;
    s.idx = idx;
    s.pack = COMPACT ? __ldg(S.site_pack + idx) : 0u;
    return s;
}

// Unpacked form of a site for the compact layout (one PRMT per field).
struct USite {
    int idx;         // dense: site index
    int x, y, z, b;  // compact: cell coordinates and basis index
};

template <bool COMPACT>
__device__ __forceinline__ USite unpack(const Site s)
{
    USite u;
    u.idx = s.idx;
    if (COMPACT) {
        u.b = (int)__byte_perm(s.pack, 0u, 0x4440);
        u.x = (int)__byte_perm(s.pack, 0u, 0x4441);
        u.y = (int)__byte_perm(s.pack, 0u, 0x4442);
        u.z = (int)__byte_perm(s.pack, 0u, 0x4443);
    } else {
        u.x = u.y = u.z = u.b = 0;
    }
    return u;
}

// P[x, y].  Dense: one 8-byte gather from the N x N array.  Compact: translation symmetry,
// P[x,y] = Pu[basis_x][(cell_y - cell_x mod size)*n_basis + basis_y]; Pu (n_basis*N*8 bytes,
// 7.2 MB at N = 30 000) stays L2-resident.  (d mod s) for d in (-s, s) is
// min_unsigned(d, d + s): no predicates.
template <bool COMPACT>
__device__ __forceinline__ double ld_pair(const SysDev &S, const USite x, const USite y)
{
    if (!COMPACT) {
        return __ldg(S.P + (long long)x.idx * S.n_sites + y.idx);
    } else {
        const int dx = y.x - x.x, dy = y.y - x.y, dz = y.z - x.z;
        const unsigned wx = min((unsigned)dx, (unsigned)(dx + S.sx));
        const unsigned wy = min((unsigned)dy, (unsigned)(dy + S.sy));
        const unsigned wz = min((unsigned)dz, (unsigned)(dz + S.sz));
        const unsigned col = ((wx * (unsigned)S.sy + wy) * (unsigned)S.sz + wz) * (unsigned)S.n_basis + (unsigned)y.b;
        const unsigned off = (unsigned)x.b * (unsigned)S.n_sites + col;
        return __ldg(S.P + off);
    }
}

template <bool COMPACT>
__device__ __forceinline__ double ld_pair(const SysDev &S, const Site x, const Site y)
{
    return ld_pair<COMPACT>(S, unpack<COMPACT>(x), unpack<COMPACT>(y));
}

template <bool COMPACT>
__device__ __forceinline__ double ld_vlat(const SysDev &S, const Site x)
{
    return __ldg(S.v_lat + ((COMPACT && !S.vlat_by_site) ? (int)(x.pack & 255u) : x.idx));
}

// np.e ** y  (core.py:2047).  The reference raises the ROUNDED constant np.e =
// 2.718281828459045 to the power y, i.e. exp(y * ln(np.e)) with ln(np.e) = 1 - 5.318e-17;
// a pow() call would recompute that logarithm for every rate.
__device__ __forceinline__ double pow_np_e(double y)
{
    const double e = exp(y);
    return fma(e, y * -5.318237706605891e-17, e);
}

// Per-process cached quantities live in shared memory for the whole launch.
struct ProcSmem {
    int *a, *b, *be;            // old site, new site, centre index of the new site
    unsigned *ap, *bp;          // packed cell/basis (compact layout)
    double *t01, *t02, *shift;  // core.py:2004-2014, 2023-2025
    double *lam, *vab, *fs;     // lambda, V_AB, 0.5*E.hop
    double *k, *cum;
};

// everything of a process except the carrier sum (new site, self term, shift, lambda...)
struct ProcStatic {
    Site a, b;
    int be;
    double t02, shift, lam, vab, fs, vl;  // vl = V_lat[b] - V_lat[a]
};

// a = current site of the carrier (index, pack, centre index e known to the caller)
template <bool COMPACT>
__device__ __forceinline__ ProcStatic load_process_static(const SysDev &S, int nn, int slot, Site a, int e,
                                                          const double *fld, int field_active)
{
    ProcStatic r;
    const long long ns = (long long)e * nn + slot;
    r.a = a;
    r.b.idx = __ldg(S.neigh + ns);
    r.b.pack = COMPACT ? __ldg(S.neigh_pack + ns) : 0u;
    const int cls = __ldg(S.site_class + a.idx);
    r.be = __ldg(S.site_centre + r.b.idx);
    r.t02 = __dmul_rn(S.qc, __dsub_rn(ld_pair<COMPACT>(S, a, a), ld_pair<COMPACT>(S, a, r.b)));
    r.shift = __dsub_rn(__ldg(S.e_rel + r.b.idx), __ldg(S.e_rel + a.idx));
    r.lam = __ldg(S.lam + cls * nn + slot);
    r.vab = __ldg(S.vab + cls * nn + slot);
    r.fs = 0.0;
    if (field_active) {
        const double *hv = S.hopvec + ns * 3;
        r.fs = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(fld[0], hv[0]), __dmul_rn(fld[1], hv[1])),
                                        __dmul_rn(fld[2], hv[2])));
    }
    r.vl = __dsub_rn(ld_vlat<COMPACT>(S, r.b), ld_vlat<COMPACT>(S, a));
    return r;
}

template <bool COMPACT>
__device__ __forceinline__ void store_process_static(ProcSmem &M, int p, const ProcStatic &r)
{
    M.a[p] = r.a.idx;
    M.b[p] = r.b.idx;
    M.be[p] = r.be;
    if (COMPACT) { M.ap[p] = r.a.pack; M.bp[p] = r.b.pack; }
    M.t02[p] = r.t02;
    M.shift[p] = r.shift;
    M.lam[p] = r.lam;
    M.vab[p] = r.vab;
    M.fs[p] = r.fs;
}

// Full re-gather of one process by one thread, carriers in order (the checker's summation
// order, so refresh_interval = 1 reproduces its rates up to exp/log rounding); loads are
// issued in batches of GB pairs to keep 2*GB gathers in flight per thread.
template <bool COMPACT>
__device__ __forceinline__ void gather_process(const SysDev &S, const int *s_occ, const unsigned *s_occp,
                                               const int *s_occe, int C, int nn, int p, const double *fld,
                                               int field_active, ProcSmem &M)
{
    constexpr int GB = 4;
    const int c = p / nn;
    Site a;
    a.idx = s_occ[c];
    a.pack = COMPACT ? s_occp[c] : 0u;
    const ProcStatic ps = load_process_static<COMPACT>(S, nn, p - c * nn, a, s_occe[c], fld, field_active);
    store_process_static<COMPACT>(M, p, ps);
    const USite ua = unpack<COMPACT>(a), ub = unpack<COMPACT>(ps.b);
    double t01 = ps.vl;
    for (int c0 = 0; c0 < C; c0 += GB) {
        double pb[GB], pa[GB];
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const int c2 = min(c0 + j, C - 1);
            Site sc;
            sc.idx = s_occ[c2];
            sc.pack = COMPACT ? s_occp[c2] : 0u;
            const USite usc = unpack<COMPACT>(sc);
            pb[j] = ld_pair<COMPACT>(S, ub, usc);
            pa[j] = ld_pair<COMPACT>(S, ua, usc);
        }
#pragma unroll
        for (int j = 0; j < GB; ++j)
            if (c0 + j < C) t01 = __dadd_rn(t01, __dmul_rn(S.qc, __dsub_rn(pb[j], pa[j])));
    }
    M.t01[p] = t01;
}

// control block written by thread 0 each step (double-buffered by step parity)
struct StepCtl {
    long long r0, r1;  // rows [r0, r1) of the time grid to fill with the post-hop displacement
    int fin;           // trajectory finished
    int pad;
};

// CT / NNT: compile-time carrier count / neighbour slots (0 = run-time values); the
// specialisation folds the shared-memory layout and the index divisions into constants.
template <int BS, bool COMPACT, int CT, int NNT>
__global__ void __launch_bounds__(BS, (BS >= 128) ? 1024 / BS : 1)
kmc_step_kernel(SysDev S_in, EnsDev E, AdvanceArgs A)
{
    const int traj = blockIdx.x;
    const int tid = threadIdx.x;
    // doped trajectory (core.py:2723-2776): its own site energies and lattice potential
    SysDev S = S_in;
    if (E.v_lat_traj) {
        S.e_rel = E.e_rel_traj + (long long)traj * S_in.n_sites;
        S.v_lat = E.v_lat_traj + (long long)traj * S_in.n_sites;
        S.vlat_by_site = 1;
    }
    const int C = CT ? CT : E.C;
    const int nn = NNT ? NNT : S.nn;
    const int n_proc = (CT && NNT) ? CT * NNT : E.n_proc;
    // fully specialised shape: one process and one (slot, carrier) item per thread
    constexpr bool FAST = (CT > 0 && NNT > 0 && CT * NNT == BS && CT % 32 == 0);
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
    double *s_terms = sd; sd += nn * C;   // moved carrier: per (slot, carrier) contributions
    double *s_wsum = sd; sd += 32;
    double *s_u = sd; sd += 4;            // [parity][u1, -log(u2)]
    StepCtl *s_ctl = reinterpret_cast<StepCtl *>(sd); sd += 2 * (sizeof(StepCtl) / sizeof(double));
    int *si = reinterpret_cast<int *>(sd);
    M.a = si; si += n_proc;
    M.b = si; si += n_proc;
    M.be = si; si += n_proc;
    M.ap = reinterpret_cast<unsigned *>(si); si += n_proc;
    M.bp = reinterpret_cast<unsigned *>(si); si += n_proc;
    int *s_occ = si; si += C;
    int *s_occe = si; si += C;
    unsigned *s_occp = reinterpret_cast<unsigned *>(si); si += C;
    int *s_wfirst = si; si += 32;   // per-warp first index with cum > u1
    int *s_sel = si; si += 2;       // tie fallback result

    if (E.done[traj]) {
        if (tid == 0 && A.steps_done) A.steps_done[traj] = 0;
        return;
    }

    const double kT = E.kT_traj ? E.kT_traj[traj] : S.kT;
    double fld[3] = {S.field[0], S.field[1], S.field[2]};
    int field_active = S.field_active;
    if (E.field_traj) {
        fld[0] = E.field_traj[3 * traj];
        fld[1] = E.field_traj[3 * traj + 1];
        fld[2] = E.field_traj[3 * traj + 2];
        field_active = (fld[0] != 0.0 || fld[1] != 0.0 || fld[2] != 0.0);
    }
    const double two_qc = __dmul_rn(2.0, S.qc);
    const long long steps_total = E.n_steps[traj];
    const unsigned long long traj_gid = E.traj_id0 + (unsigned long long)traj;
    const int R = E.refresh_interval;
    // the thread that prepares the next step's draws while the others wait on gathers
    const int rng_tid = (BS > 32) ? 32 : 0;

    auto draw = [&](long long step_local, double *dst) {  // u1 and -log(u2) of a step
        double u1, u2;
        if (E.rng_mode == PYCD_RNG_REPLAY) {
            if (step_local < A.max_steps) {
                const double *dr = A.draws + ((long long)traj * A.max_steps + step_local) * 2;
                u1 = dr[0];
                u2 = dr[1];
            } else {
                u1 = 0.0; u2 = 1.0;
            }
        } else {
            philox_uniforms(E.seed, traj_gid, (unsigned long long)(steps_total + step_local), u1, u2);
        }
        dst[0] = u1;
        dst[1] = -log(u2);
    };

    for (int c = tid; c < C; c += BS) {
        const int s = E.occ[(long long)traj * C + c];
        s_occ[c] = s;
        s_occe[c] = S.site_centre[s];
        s_occp[c] = COMPACT ? S.site_pack[s] : 0u;
    }
    for (int d = tid; d < 3 * C; d += BS) {
        s_disp[d] = E.disp[(long long)traj * 3 * C + d];
        s_row[d] = E.row[(long long)traj * 3 * C + d];
        s_drift[d] = E.drift[(long long)traj * 3 * C + d];
    }
    // thread-0 scalars
    double t = E.t[traj];
    double energy = (E.energy && tid == 0) ? E.energy[traj] : 0.0;
    long long start = E.start_idx[traj];
    long long n_tie = 0, n_clamp = 0;
    long long step_local = 0;
    int finished = 0;
    // steps until the next full re-gather (launches start on a multiple of R)
    int to_refresh = (R <= 1) ? 0 : (int)((R - (steps_total % R)) % R);
    if (tid == 0) {
        s_ctl[0].r0 = s_ctl[0].r1 = 0; s_ctl[0].fin = 0;
        s_ctl[1].r0 = s_ctl[1].r1 = 0; s_ctl[1].fin = 0;
    }
    if (tid == rng_tid) draw(0, s_u);
    __syncthreads();

    while (true) {
        const int par = (int)(step_local & 1);
        // ---- recording decided by the previous step: rows [r0, r1) ----
        {
            const StepCtl ctl = s_ctl[par ^ 1];
            if (ctl.r1 > ctl.r0) {
                // unwrapped[start:end] = unwrapped[start-1] + displacement, core.py:2852-2854
                for (int d = tid; d < 3 * C; d += BS) {
                    const double v = s_row[d] + s_disp[d];
                    s_row[d] = v;
                    s_disp[d] = 0.0;
                    if (E.unwrapped) {
                        double *dst = E.unwrapped + ((long long)traj * E.n_path + ctl.r0) * 3 * C + d;
                        for (long long r = ctl.r0; r < ctl.r1; ++r, dst += 3 * C) *dst = v;
                    }
                }
            }
            finished = ctl.fin;
        }
        if (finished || step_local >= A.max_steps) break;

        // ---- rates + inclusive scan of the raw rates (BS-wide groups) ----
        const bool full = (to_refresh == 0);
        to_refresh = (R <= 1) ? 0 : (full ? R - 1 : to_refresh - 1);
        double carry = 0.0;
        for (int base = 0; base < n_proc; base += BS) {
            const int p = base + tid;
            double kp = 0.0;
            if (p < n_proc) {
                if (full) gather_process<COMPACT>(S, s_occ, s_occp, s_occe, C, nn, p, fld, field_active, M);
                const double ew = __dmul_rn(two_qc, __dadd_rn(M.t01[p], M.t02[p]));     // core.py:2016
                const double g0 = __dadd_rn(ew, M.shift[p]);
                const double lam = M.lam[p];
                const double lg = __dadd_rn(lam, g0);
                const double gs = __dsub_rn(__dsub_rn(__ddiv_rn(__dmul_rn(lg, lg), __dmul_rn(4.0, lam)),
                                                      M.vab[p]), M.fs[p]);               // core.py:2045
                kp = __dmul_rn(S.vn, pow_np_e(__ddiv_rn(-gs, kT)));                       // core.py:2047
                M.k[p] = kp;
            }
            double x = kp;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wsum[wid] = x;
            __syncthreads();  // (A)
            double pre = carry, tot = carry;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const double v = s_wsum[w];
                if (w < wid) pre += v;
                tot += v;
            }
            if (p < n_proc) M.cum[p] = pre + x;   // un-normalised inclusive prefix
            carry = tot;
            if (base + BS < n_proc) __syncthreads();  // s_wsum reused by the next group
        }
        const double ktot = carry;
        const double u1 = s_u[2 * par], nlog_u2 = s_u[2 * par + 1];
        // cum/k_total > u1  <=>  cum > u1*k_total up to rounding; draws within TIE_TOL of an
        // edge are re-decided below in the reference's sequential order
        const double thresh = u1 * ktot, tie_w = TIE_TOL * ktot;

        // ---- first index with cumsum(k/k_total) > u1, core.py:2797-2800 ----
        int first = INT_MAX;
        for (int base = 0; base < n_proc; base += BS) {
            const int p = base + tid;
            const bool hit = (p < n_proc) && (M.cum[p] > thresh);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m && first == INT_MAX) first = base + wid * 32 + (__ffs(m) - 1);
        }
        if (lane == 0) s_wfirst[wid] = first;
        __syncthreads();  // (B) M.cum, s_wfirst visible
        int sel = INT_MAX;
#pragma unroll
        for (int w = 0; w < NW; ++w) sel = min(sel, s_wfirst[w]);
        bool tie = (sel == INT_MAX);
        if (!tie) {
            const double hi = M.cum[sel], lo = sel > 0 ? M.cum[sel - 1] : 0.0;
            tie = (hi - thresh < tie_w) || (sel > 0 && thresh - lo < tie_w);
        }
        if (tie) {  // block-uniform: redo the selection in the reference's sequential order
            if (tid == 0) {
                double kseq = 0.0;
                for (int p = 0; p < n_proc; ++p) kseq += M.k[p];
                double cum = 0.0;
                int s2 = -1;
                for (int p = 0; p < n_proc; ++p) {
                    cum += M.k[p] / kseq;
                    if (cum > u1) { s2 = p; break; }
                }
                if (s2 < 0) { s2 = n_proc - 1; ++n_clamp; }  // the reference raises IndexError here
                ++n_tie;
                s_sel[0] = s2;
            }
            __syncthreads();
            sel = s_sel[0];
        }

        const int cs = sel / nn, slot = sel - cs * nn;
        Site a_old, b_new;
        a_old.idx = M.a[sel];
        b_new.idx = M.b[sel];
        a_old.pack = COMPACT ? M.ap[sel] : 0u;
        b_new.pack = COMPACT ? M.bp[sel] : 0u;
        const int e_new = M.be[sel];
        const USite u_old = unpack<COMPACT>(a_old), u_new = unpack<COMPACT>(b_new);
        const bool next_full = (to_refresh == 0);

        // ---- issue the long-latency loads of this step's tail first ----
        double hv0 = 0.0, hv1 = 0.0, hv2 = 0.0;
        if (tid == 0) {  // hop vector of the selected process
            const double *hv = S.hopvec + ((long long)s_occe[cs] * nn + slot) * 3;
            hv0 = __ldg(hv); hv1 = __ldg(hv + 1); hv2 = __ldg(hv + 2);
        }
        double patch[4];
        int npatch = 0;
        ProcStatic moved;
        if (!next_full) {
            if constexpr (FAST) {
                // one process and one (slot, carrier) item per thread; the moved carrier's own
                // threads have no patch and prefetch their new process instead
                if (tid / NNT != cs) {
                    Site pa, pb;
                    pa.idx = M.a[tid]; pb.idx = M.b[tid];
                    pa.pack = COMPACT ? M.ap[tid] : 0u;
                    pb.pack = COMPACT ? M.bp[tid] : 0u;
                    const USite ua = unpack<COMPACT>(pa), ub = unpack<COMPACT>(pb);
                    const double nb_ = ld_pair<COMPACT>(S, ub, u_new), na_ = ld_pair<COMPACT>(S, ua, u_new);
                    const double ob_ = ld_pair<COMPACT>(S, ub, u_old), oa_ = ld_pair<COMPACT>(S, ua, u_old);
                    patch[0] = S.qc * (nb_ - na_) - S.qc * (ob_ - oa_);
                } else {
                    moved = load_process_static<COMPACT>(S, nn, tid % NNT, b_new, e_new, fld, field_active);
                }
                const int sl = tid / CT, c2 = tid % CT;   // warp-uniform slot (CT % 32 == 0)
                Site nbr;
                nbr.idx = __ldg(S.neigh + (long long)e_new * nn + sl);
                nbr.pack = COMPACT ? __ldg(S.neigh_pack + (long long)e_new * nn + sl) : 0u;
                Site sc;
                if (c2 == cs) sc = b_new;
                else { sc.idx = s_occ[c2]; sc.pack = COMPACT ? s_occp[c2] : 0u; }
                const USite usc = unpack<COMPACT>(sc);
                double term = S.qc * (ld_pair<COMPACT>(S, unpack<COMPACT>(nbr), usc) - ld_pair<COMPACT>(S, u_new, usc));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
                if (lane == 0) s_terms[wid] = term;   // CT/32 partial sums per slot
            } else {
                // (a) untouched processes: 4 elements each
                for (int p = tid; p < n_proc && npatch < 4; p += BS) {
                    double v = 0.0;
                    if (p / nn != cs) {
                        Site pa, pb;
                        pa.idx = M.a[p]; pb.idx = M.b[p];
                        pa.pack = COMPACT ? M.ap[p] : 0u;
                        pb.pack = COMPACT ? M.bp[p] : 0u;
                        const USite ua = unpack<COMPACT>(pa), ub = unpack<COMPACT>(pb);
                        const double nb_ = ld_pair<COMPACT>(S, ub, u_new), na_ = ld_pair<COMPACT>(S, ua, u_new);
                        const double ob_ = ld_pair<COMPACT>(S, ub, u_old), oa_ = ld_pair<COMPACT>(S, ua, u_old);
                        v = S.qc * (nb_ - na_) - S.qc * (ob_ - oa_);
                    }
                    patch[npatch++] = v;
                }
                // (b) the moved carrier's nn processes: (slot, carrier) items over the whole block
                for (int item = tid; item < nn * C; item += BS) {
                    const int sl = item / C, c2 = item - sl * C;
                    Site nbr;
                    nbr.idx = __ldg(S.neigh + (long long)e_new * nn + sl);
                    nbr.pack = COMPACT ? __ldg(S.neigh_pack + (long long)e_new * nn + sl) : 0u;
                    Site sc;
                    if (c2 == cs) sc = b_new;
                    else { sc.idx = s_occ[c2]; sc.pack = COMPACT ? s_occp[c2] : 0u; }
                    const USite usc = unpack<COMPACT>(sc);
                    s_terms[item] = S.qc * (ld_pair<COMPACT>(S, unpack<COMPACT>(nbr), usc) -
                                            ld_pair<COMPACT>(S, u_new, usc));
                }
                // (c) static parts of the moved carrier's new processes: lane j of warp w <-> slot w + j*NW
                const int my_slot = wid + lane * NW;
                if (my_slot < nn)
                    moved = load_process_static<COMPACT>(S, nn, my_slot, b_new, e_new, fld, field_active);
            }
        }

        // ---- thread 0: time advance, grid bookkeeping, hop, core.py:2802-2830, 2844-2861 ----
        if (tid == 0) {
            t += nlog_u2 / ktot;
            const long long end = (long long)(t / (E.dt_grid_traj ? E.dt_grid_traj[traj] : E.dt_grid));
            const long long start_before = start;
            StepCtl ctl;
            ctl.r0 = 0; ctl.r1 = 0; ctl.fin = 0; ctl.pad = 0;
            if (end >= start + 1) {
                const long long e2 = end >= E.n_path ? E.n_path : end;
                if (start < E.n_path) { ctl.r0 = start; ctl.r1 = e2; }
                start = e2;
            }
            if (E.energy) {  // output_data energy / delg_0, core.py:2807-2809, 2826, 2855-2857
                const double g0 = __dadd_rn(__dmul_rn(two_qc, __dadd_rn(M.t01[sel], M.t02[sel])), M.shift[sel]);
                const long long lo_r = start_before, hi_r = end < E.n_path ? end : E.n_path;
                for (long long r = lo_r; r < hi_r; ++r) E.dg0_grid[(long long)traj * E.n_path + r] = g0;
                energy += g0;
                for (long long r = ctl.r0; r < ctl.r1; ++r) E.energy_grid[(long long)traj * E.n_path + r] = energy;
            }
            if (E.stop_at_grid_end && end >= E.n_path) ctl.fin = 1;
            if (E.step_limit > 0 && steps_total + step_local + 1 >= E.step_limit) ctl.fin = 1;
            s_ctl[par] = ctl;
            // the other threads only read the entries of the carriers that did not move
            s_occ[cs] = b_new.idx;
            s_occe[cs] = e_new;
            if (COMPACT) s_occp[cs] = b_new.pack;
            const double kp = M.k[sel];
            s_disp[3 * cs] += hv0; s_disp[3 * cs + 1] += hv1; s_disp[3 * cs + 2] += hv2;
            if (field_active) {
                s_drift[3 * cs] += hv0 * kp; s_drift[3 * cs + 1] += hv1 * kp; s_drift[3 * cs + 2] += hv2 * kp;
            }
            if (A.events_out) A.events_out[(long long)traj * A.max_steps + step_local] = sel;
            if (A.times_out) A.times_out[(long long)traj * A.max_steps + step_local] = t;
        }
        // ---- next step's draws (off the critical path: the block is waiting on gathers) ----
        if (tid == rng_tid) draw(step_local + 1, s_u + 2 * (par ^ 1));
        __syncthreads();  // (C) s_terms, s_ctl, s_occ, s_disp visible; all reads of M.*[sel] done

        // ---- bring the cached sums up to date for the next step ----
        if (!next_full) {
            if constexpr (FAST) {
                if (tid / NNT != cs) {
                    M.t01[tid] += patch[0];
                } else {
                    constexpr int PARTS = CT / 32;
                    const int sl = tid % NNT;
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < PARTS; ++j) acc += s_terms[sl * PARTS + j];
                    store_process_static<COMPACT>(M, tid, moved);
                    M.t01[tid] = moved.vl + acc;
                }
            } else {
                int ip = 0;
                for (int p = tid; p < n_proc; p += BS, ++ip) {
                    if (p / nn == cs) continue;
                    if (ip < 4) {
                        M.t01[p] += patch[ip];
                    } else {
                        Site pa, pb;
                        pa.idx = M.a[p]; pb.idx = M.b[p];
                        pa.pack = COMPACT ? M.ap[p] : 0u;
                        pb.pack = COMPACT ? M.bp[p] : 0u;
                        const USite ua = unpack<COMPACT>(pa), ub = unpack<COMPACT>(pb);
                        M.t01[p] += S.qc * (ld_pair<COMPACT>(S, ub, u_new) - ld_pair<COMPACT>(S, ua, u_new)) -
                                    S.qc * (ld_pair<COMPACT>(S, ub, u_old) - ld_pair<COMPACT>(S, ua, u_old));
                    }
                }
                // one warp per slot of the moved carrier: reduce its C contributions
                int j = 0;
                for (int sl = wid; sl < nn; sl += NW, ++j) {
                    double acc = 0.0;
                    for (int c2 = lane; c2 < C; c2 += 32) acc += s_terms[sl * C + c2];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == j) {
                        const int p = cs * nn + sl;
                        store_process_static<COMPACT>(M, p, moved);
                        M.t01[p] = moved.vl + acc;
                    }
                }
            }
        }
        ++step_local;
        // (D) only the generic incremental path hands entries of M.* to non-owner threads
        if (!FAST && !next_full) __syncthreads();
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
        if (E.energy) E.energy[traj] = energy;
        E.start_idx[traj] = start;
        E.n_steps[traj] = steps_total + step_local;
        E.near_tie[traj] += n_tie;
        E.clamped[traj] += n_clamp;
        if (finished) E.done[traj] = 1;
        if (A.steps_done) A.steps_done[traj] = step_local;
    }
}

// ---------------------------------------------------------------------------------------
// Carrier-per-thread variant: thread c owns carrier c and its NN processes (block = CT
// threads, CT = number of carriers, a multiple of 32).  The per-process state lives in
// registers, the old-site elements P[a, .] are shared by the NN slots of a carrier, the scan is
// one local prefix + one warp scan, and a step needs 3 block barriers.  Same arithmetic as
// kmc_step_kernel (stateless order for R = 1, incremental patches otherwise); ~3x fewer warp
// instructions per KMC step at (CT, NN) = (64, 4).
// LEAN: <= 128 registers (8 CTAs of 64 threads per SM) for ensembles larger than the ~4 CTAs/SM
// the register-rich variant keeps resident; the periodic re-gather then uses smaller batches.
template <int CT, int NN, bool COMPACT, bool LEAN>
__global__ void __launch_bounds__(CT, LEAN ? (512 / CT) : 1)
kmc_step_carrier_kernel(SysDev S, EnsDev E, AdvanceArgs A)
{
    constexpr int NW = CT / 32;
    constexpr int NP = CT * NN;
    static_assert(CT % 32 == 0 && CT >= 32, "one warp-aligned thread per carrier");
    const int traj = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;

    __shared__ double s_k[NP], s_cum[NP];          // rates and running sums of this step
    __shared__ int s_b[NP], s_be[NP];              // new site / its centre index per process
    __shared__ unsigned s_bp[NP];
    __shared__ int s_nb[NP][NN];                   // neighbours of each process's new site (and packs):
    __shared__ unsigned s_nbp[NP][NN];             // after a hop the next gathers need no index load
    __shared__ int s_occ[CT], s_occe[CT];
    __shared__ unsigned s_occp[CT];
    __shared__ double s_disp[3 * CT], s_row[3 * CT], s_drift[3 * CT];
    __shared__ double s_red[NN][NW > 1 ? NW : 1];
    __shared__ double s_wsum[NW > 1 ? NW : 1];
    __shared__ double s_u[4];
    __shared__ StepCtl s_ctl[2];
    __shared__ int s_wfirst[NW > 1 ? NW : 1];
    __shared__ int s_sel[2];
    __shared__ double s_g0[NP];                    // delta-G0 per process (energy outputs only)

    if (E.done[traj]) {
        if (tid == 0 && A.steps_done) A.steps_done[traj] = 0;
        return;
    }
    const double kT = E.kT_traj ? E.kT_traj[traj] : S.kT;
    double fld[3] = {S.field[0], S.field[1], S.field[2]};
    int field_active = S.field_active;
    if (E.field_traj) {
        fld[0] = E.field_traj[3 * traj];
        fld[1] = E.field_traj[3 * traj + 1];
        fld[2] = E.field_traj[3 * traj + 2];
        field_active = (fld[0] != 0.0 || fld[1] != 0.0 || fld[2] != 0.0);
    }
    const double two_qc = __dmul_rn(2.0, S.qc);
    const long long steps_total = E.n_steps[traj];
    const unsigned long long traj_gid = E.traj_id0 + (unsigned long long)traj;
    const int R = E.refresh_interval;
    const int rng_tid = (CT > 32) ? 32 : 0;
    const bool want_energy = (E.energy != nullptr);

    auto draw = [&](long long step_local, double *dst) {  // u1 and -log(u2) of a step
        double u1, u2;
        if (E.rng_mode == PYCD_RNG_REPLAY) {
            if (step_local < A.max_steps) {
                const double *dr = A.draws + ((long long)traj * A.max_steps + step_local) * 2;
                u1 = dr[0];
                u2 = dr[1];
            } else {
                u1 = 0.0; u2 = 1.0;
            }
        } else {
            philox_uniforms(E.seed, traj_gid, (unsigned long long)(steps_total + step_local), u1, u2);
        }
        dst[0] = u1;
        dst[1] = -log(u2);
    };

    // ---- per-thread (per-carrier) state in registers ----
    Site a;            // current site of this carrier
    int ae;            // its centre index
    Site b[NN];
    int be[NN];
    double t01[NN], t02[NN], shift[NN], lam[NN], vab[NN], fs[NN], i4l[NN];
    int nb2[NN][NN];
    unsigned nb2p[NN][NN];
    const double neg_inv_kT = -1.0 / kT;

    // everything of the NN processes but the carrier sums.  nbr: the new sites if already known
    // (cached neighbour list of the selected process), else they are loaded from the table.
    auto load_static = [&](Site na, int ne, const Site *nbr) {
        a = na; ae = ne;
        const int cls = __ldg(S.site_class + na.idx);
        const USite ua = unpack<COMPACT>(na);
        const double paa = ld_pair<COMPACT>(S, ua, ua), era = __ldg(S.e_rel + na.idx);
        const double vla = ld_vlat<COMPACT>(S, na);
#pragma unroll
        for (int s = 0; s < NN; ++s) {
            const long long ns = (long long)ne * NN + s;
            if (nbr) {
                b[s] = nbr[s];
            } else {
                b[s].idx = __ldg(S.neigh + ns);
                b[s].pack = COMPACT ? __ldg(S.neigh_pack + ns) : 0u;
            }
            be[s] = __ldg(S.site_centre + b[s].idx);
            t02[s] = __dmul_rn(S.qc, __dsub_rn(paa, ld_pair<COMPACT>(S, ua, unpack<COMPACT>(b[s]))));
            shift[s] = __dsub_rn(__ldg(S.e_rel + b[s].idx), era);
            lam[s] = __ldg(S.lam + cls * NN + s);
            vab[s] = __ldg(S.vab + cls * NN + s);
            i4l[s] = __ldg(S.i4l + cls * NN + s);
            fs[s] = 0.0;
            if (field_active) {
                const double *hv = S.hopvec + ns * 3;
                fs[s] = __dmul_rn(0.5, __dadd_rn(__dadd_rn(__dmul_rn(fld[0], hv[0]), __dmul_rn(fld[1], hv[1])),
                                                 __dmul_rn(fld[2], hv[2])));
            }
            t01[s] = __dsub_rn(ld_vlat<COMPACT>(S, b[s]), vla);
#pragma unroll
            for (int s2 = 0; s2 < NN; ++s2) {
                nb2[s][s2] = __ldg(S.neigh2 + ns * NN + s2);
                nb2p[s][s2] = COMPACT ? __ldg(S.neigh2_pack + ns * NN + s2) : 0u;
            }
        }
    };
    auto publish = [&]() {                      // make the new sites of my processes visible
#pragma unroll
        for (int s = 0; s < NN; ++s) {
            s_b[tid * NN + s] = b[s].idx;
            s_bp[tid * NN + s] = b[s].pack;
            s_be[tid * NN + s] = be[s];
#pragma unroll
            for (int s2 = 0; s2 < NN; ++s2) {
                s_nb[tid * NN + s][s2] = nb2[s][s2];
                s_nbp[tid * NN + s][s2] = nb2p[s][s2];
            }
        }
    };

    {
        const int s0 = E.occ[(long long)traj * CT + tid];
        s_occ[tid] = s0;
        s_occe[tid] = S.site_centre[s0];
        s_occp[tid] = COMPACT ? S.site_pack[s0] : 0u;
    }
    for (int d = tid; d < 3 * CT; d += CT) {
        s_disp[d] = E.disp[(long long)traj * 3 * CT + d];
        s_row[d] = E.row[(long long)traj * 3 * CT + d];
        s_drift[d] = E.drift[(long long)traj * 3 * CT + d];
    }
    double t = E.t[traj];
    double energy = (E.energy && tid == 0) ? E.energy[traj] : 0.0;
    long long start = E.start_idx[traj];
    long long n_tie = 0, n_clamp = 0;
    long long step_local = 0;
    int finished = 0;
    int to_refresh = (R <= 1) ? 0 : (int)((R - (steps_total % R)) % R);
    if (tid == 0) {
        s_ctl[0].r0 = s_ctl[0].r1 = 0; s_ctl[0].fin = 0;
        s_ctl[1].r0 = s_ctl[1].r1 = 0; s_ctl[1].fin = 0;
    }
    if (tid == rng_tid) draw(0, s_u);
    __syncthreads();
    {
        Site na;
        na.idx = s_occ[tid];
        na.pack = s_occp[tid];
        load_static(na, s_occe[tid], nullptr);
        publish();
    }
    bool need_full = true;   // the cached sums are rebuilt at the first step of every launch
    __syncthreads();

    while (true) {
        const int par = (int)(step_local & 1);
        {
            const StepCtl ctl = s_ctl[par ^ 1];
            if (ctl.r1 > ctl.r0) {  // unwrapped[start:end] = unwrapped[start-1] + displacement, core.py:2852-2854
                for (int d = tid; d < 3 * CT; d += CT) {
                    const double v = s_row[d] + s_disp[d];
                    s_row[d] = v;
                    s_disp[d] = 0.0;
                    if (E.unwrapped) {
                        double *dst = E.unwrapped + ((long long)traj * E.n_path + ctl.r0) * 3 * CT + d;
                        for (long long r = ctl.r0; r < ctl.r1; ++r, dst += 3 * CT) *dst = v;
                    }
                }
            }
            finished = ctl.fin;
        }
        if (finished || step_local >= A.max_steps) break;

        // ---- full re-gather (every R steps; every step for R = 1): carriers in order ----
        const bool full = need_full || (to_refresh == 0);
        need_full = false;
        to_refresh = (R <= 1) ? 0 : ((to_refresh == 0) ? R - 1 : to_refresh - 1);
        if (full) {
            const USite ua = unpack<COMPACT>(a);
            USite ub[NN];
#pragma unroll
            for (int s = 0; s < NN; ++s) {
                ub[s] = unpack<COMPACT>(b[s]);
                t01[s] = __dsub_rn(ld_vlat<COMPACT>(S, b[s]), ld_vlat<COMPACT>(S, a));
            }
            constexpr int GB = LEAN ? 2 : 8;
            for (int c0 = 0; c0 < CT; c0 += GB) {
                double pa[GB], pb[GB][NN];
#pragma unroll
                for (int j = 0; j < GB; ++j) {
                    Site sc;
                    sc.idx = s_occ[c0 + j];
                    sc.pack = s_occp[c0 + j];
                    const USite usc = unpack<COMPACT>(sc);
                    pa[j] = ld_pair<COMPACT>(S, ua, usc);
#pragma unroll
                    for (int s = 0; s < NN; ++s) pb[j][s] = ld_pair<COMPACT>(S, ub[s], usc);
                }
#pragma unroll
                for (int j = 0; j < GB; ++j)
#pragma unroll
                    for (int s = 0; s < NN; ++s)
                        t01[s] = __dadd_rn(t01[s], __dmul_rn(S.qc, __dsub_rn(pb[j][s], pa[j])));
            }
        }

        // ---- rates, local prefix ----
        double k[NN], loc[NN];
        double run = 0.0;
#pragma unroll
        for (int s = 0; s < NN; ++s) {
            const double ew = __dmul_rn(two_qc, __dadd_rn(t01[s], t02[s]));              // core.py:2016
            const double g0 = __dadd_rn(ew, shift[s]);
            const double lg = __dadd_rn(lam[s], g0);
            if (R <= 1) {   // stateless mode: the reference's operation order, divisions included
                const double gs = __dsub_rn(__dsub_rn(__ddiv_rn(__dmul_rn(lg, lg), __dmul_rn(4.0, lam[s])),
                                                      vab[s]), fs[s]);                       // core.py:2045
                k[s] = __dmul_rn(S.vn, pow_np_e(__ddiv_rn(-gs, kT)));                         // core.py:2047
            } else {        // incremental mode: cached reciprocals (<= 1e-14 relative in the rate)
                const double gs = (lg * lg) * i4l[s] - vab[s] - fs[s];
                k[s] = S.vn * pow_np_e(gs * neg_inv_kT);
            }
            run += k[s];
            loc[s] = run;
            s_k[tid * NN + s] = k[s];
            if (want_energy) s_g0[tid * NN + s] = g0;
        }
        // ---- warp scan of the per-thread totals, cross-warp prefix ----
        double x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        double pre = x - run, ktot;   // exclusive prefix inside the warp
        if (NW > 1) {
            if (lane == 31) s_wsum[wid] = x;
            __syncthreads();  // (A)
            double before = 0.0, tot = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const double v = s_wsum[w];
                if (w < wid) before += v;
                tot += v;
            }
            pre += before;
            ktot = tot;
        } else {
            ktot = __shfl_sync(0xffffffffu, x, 31);
        }
        const double u1 = s_u[2 * par], nlog_u2 = s_u[2 * par + 1];
        const double thresh = u1 * ktot, tie_w = TIE_TOL * ktot;
        int first_local = NN;
#pragma unroll
        for (int s = NN - 1; s >= 0; --s) {
            const double cum = pre + loc[s];
            s_cum[tid * NN + s] = cum;
            if (cum > thresh) first_local = s;
        }
        const unsigned m = __ballot_sync(0xffffffffu, first_local < NN);
        int first = INT_MAX;
        if (m) {
            const int src = __ffs(m) - 1;
            first = (wid * 32 + src) * NN + __shfl_sync(0xffffffffu, first_local, src);
        }
        if (NW > 1) {
            if (lane == 0) s_wfirst[wid] = first;
        }
        __syncthreads();  // (B) s_k, s_cum, s_wfirst visible
        int sel = first;
        if (NW > 1) {
            sel = INT_MAX;
#pragma unroll
            for (int w = 0; w < NW; ++w) sel = min(sel, s_wfirst[w]);
        }
        bool tie = (sel == INT_MAX);
        if (!tie) {
            const double hi = s_cum[sel], lo = sel > 0 ? s_cum[sel - 1] : 0.0;
            tie = (hi - thresh < tie_w) || (sel > 0 && thresh - lo < tie_w);
        }
        if (tie) {  // block-uniform: redo the selection in the reference's sequential order
            if (tid == 0) {
                double kseq = 0.0;
                for (int p = 0; p < NP; ++p) kseq += s_k[p];
                double cum = 0.0;
                int s2 = -1;
                for (int p = 0; p < NP; ++p) {
                    cum += s_k[p] / kseq;
                    if (cum > u1) { s2 = p; break; }
                }
                if (s2 < 0) { s2 = NP - 1; ++n_clamp; }
                ++n_tie;
                s_sel[0] = s2;
            }
            __syncthreads();
            sel = s_sel[0];
        }

        const int cs = sel / NN, slot = sel - cs * NN;
        Site a_old, b_new;
        a_old.idx = s_occ[cs];
        a_old.pack = s_occp[cs];
        b_new.idx = s_b[sel];
        b_new.pack = s_bp[sel];
        const int e_old = s_occe[cs], e_new = s_be[sel];
        const USite u_old = unpack<COMPACT>(a_old), u_new = unpack<COMPACT>(b_new);
        const bool next_full = (to_refresh == 0);

        // ---- long-latency loads of the tail, issued before the barrier ----
        double hv0 = 0.0, hv1 = 0.0, hv2 = 0.0;
        if (tid == 0) {
            const double *hv = S.hopvec + ((long long)e_old * NN + slot) * 3;
            hv0 = __ldg(hv); hv1 = __ldg(hv + 1); hv2 = __ldg(hv + 2);
        }
        double patch[NN];
        if (!next_full) {
            // (1) issue every gather first, (2) then consume: one L2 round trip for the whole tail
            Site sc = (tid == cs) ? b_new : Site{s_occ[tid], s_occp[tid]};
            const USite usc = unpack<COMPACT>(sc);
            const double p_base = ld_pair<COMPACT>(S, u_new, usc);
            double p_nbr[NN];
#pragma unroll
            for (int s = 0; s < NN; ++s) {
                Site nbr;   // new site of slot s of the moved carrier: cached with the selected process
                nbr.idx = s_nb[sel][s];
                nbr.pack = s_nbp[sel][s];
                p_nbr[s] = ld_pair<COMPACT>(S, unpack<COMPACT>(nbr), usc);
            }
            if (tid != cs) {
                // untouched carrier: q_c [(P[b_s,b] - P[a,b]) - (P[b_s,a_old] - P[a,a_old])] per slot
                const USite ua = unpack<COMPACT>(a);
                const double pa_new = ld_pair<COMPACT>(S, ua, u_new), pa_old = ld_pair<COMPACT>(S, ua, u_old);
                double pb_new[NN], pb_old[NN];
#pragma unroll
                for (int s = 0; s < NN; ++s) {
                    const USite ub = unpack<COMPACT>(b[s]);
                    pb_new[s] = ld_pair<COMPACT>(S, ub, u_new);
                    pb_old[s] = ld_pair<COMPACT>(S, ub, u_old);
                }
#pragma unroll
                for (int s = 0; s < NN; ++s)
                    patch[s] = S.qc * (pb_new[s] - pa_new) - S.qc * (pb_old[s] - pa_old);
            } else {   // my carrier moved: new processes (t01 = V_lat part for now)
                Site nbr[NN];
#pragma unroll
                for (int s = 0; s < NN; ++s) { nbr[s].idx = s_nb[sel][s]; nbr[s].pack = s_nbp[sel][s]; }
                load_static(b_new, e_new, nbr);
            }
            // contribution of MY carrier's site to the moved carrier's new processes
            double term[NN];
#pragma unroll
            for (int s = 0; s < NN; ++s) term[s] = S.qc * (p_nbr[s] - p_base);
#pragma unroll
            for (int s = 0; s < NN; ++s) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) term[s] += __shfl_xor_sync(0xffffffffu, term[s], o);
                if (lane == 0) s_red[s][wid] = term[s];
            }
        }

        // ---- thread 0: time advance, grid bookkeeping, hop, core.py:2802-2830, 2844-2861 ----
        if (tid == 0) {
            t += nlog_u2 / ktot;
            const long long end = (long long)(t / (E.dt_grid_traj ? E.dt_grid_traj[traj] : E.dt_grid));
            const long long start_before = start;
            StepCtl ctl;
            ctl.r0 = 0; ctl.r1 = 0; ctl.fin = 0; ctl.pad = 0;
            if (end >= start + 1) {
                const long long e2 = end >= E.n_path ? E.n_path : end;
                if (start < E.n_path) { ctl.r0 = start; ctl.r1 = e2; }
                start = e2;
            }
            if (E.energy) {  // output_data energy / delg_0, core.py:2807-2809, 2826, 2855-2857
                const double g0 = s_g0[sel];
                const long long hi_r = end < E.n_path ? end : E.n_path;
                for (long long r = start_before; r < hi_r; ++r) E.dg0_grid[(long long)traj * E.n_path + r] = g0;
                energy += g0;
                for (long long r = ctl.r0; r < ctl.r1; ++r) E.energy_grid[(long long)traj * E.n_path + r] = energy;
            }
            if (E.stop_at_grid_end && end >= E.n_path) ctl.fin = 1;
            if (E.step_limit > 0 && steps_total + step_local + 1 >= E.step_limit) ctl.fin = 1;
            s_ctl[par] = ctl;
            const double kp = s_k[sel];
            s_disp[3 * cs] += hv0; s_disp[3 * cs + 1] += hv1; s_disp[3 * cs + 2] += hv2;
            if (field_active) {
                s_drift[3 * cs] += hv0 * kp; s_drift[3 * cs + 1] += hv1 * kp; s_drift[3 * cs + 2] += hv2 * kp;
            }
            if (A.events_out) A.events_out[(long long)traj * A.max_steps + step_local] = sel;
            if (A.times_out) A.times_out[(long long)traj * A.max_steps + step_local] = t;
        }
        if (tid == rng_tid) draw(step_local + 1, s_u + 2 * (par ^ 1));
        __syncthreads();  // (C) s_red, s_ctl, s_disp visible; all reads of s_occ[cs] / s_b[sel] done

        if (tid == cs) {
            s_occ[cs] = b_new.idx;
            s_occe[cs] = e_new;
            s_occp[cs] = b_new.pack;
            if (next_full) load_static(b_new, e_new, nullptr);
            publish();
        }
        if (!next_full) {
            if (tid != cs) {
#pragma unroll
                for (int s = 0; s < NN; ++s) t01[s] += patch[s];
            } else {
#pragma unroll
                for (int s = 0; s < NN; ++s) {
                    double acc = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) acc += s_red[s][w];
                    t01[s] += acc;
                }
            }
        }
        ++step_local;
        // s_occ / s_b of the moved carrier are read by the others only after barriers (A)/(B) of the
        // next step -- except by a full re-gather, which starts right away
        if (next_full) __syncthreads();
    }

    // ---- write the state back ----
    __syncthreads();
    E.occ[(long long)traj * CT + tid] = s_occ[tid];
    for (int d = tid; d < 3 * CT; d += CT) {
        E.disp[(long long)traj * 3 * CT + d] = s_disp[d];
        E.row[(long long)traj * 3 * CT + d] = s_row[d];
        E.drift[(long long)traj * 3 * CT + d] = s_drift[d];
    }
    if (step_local > 0)
        for (int p = tid; p < NP; p += CT) E.rates[(long long)traj * NP + p] = s_k[p];
    if (tid == 0) {
        E.t[traj] = t;
        if (E.energy) E.energy[traj] = energy;
        E.start_idx[traj] = start;
        E.n_steps[traj] = steps_total + step_local;
        E.near_tie[traj] += n_tie;
        E.clamped[traj] += n_clamp;
        if (finished) E.done[traj] = 1;
        if (A.steps_done) A.steps_done[traj] = step_local;
    }
}

// V_lat = P . q_lat, one warp per row, double-double accumulation so that the
// differences V_lat[b]-V_lat[a] keep ~1e-16 Ha accuracy at N = 30 000.
__global__ void __launch_bounds__(256)
vlat_kernel(const double *__restrict__ P, const double *__restrict__ q, long long n_rows, long long n,
            double *__restrict__ v)
{
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
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

static size_t kmc_smem_bytes(int n_proc, int C, int nn) {
    const size_t doubles = (size_t)8 * n_proc + (size_t)9 * C + (size_t)nn * C + 32 + 4 + 2 * 3;
    const size_t ints = (size_t)5 * n_proc + 3 * (size_t)C + 32 + 2;
    return doubles * 8 + ((ints + 1) / 2) * 8;
}

// basis | x<<8 | y<<16 | z<<24 per site (site = cell*n_basis + basis, cell = (x*sy + y)*sz + z)
// Lattice potential of doped trajectories: the dopant sites carry the dopant's charge instead of the
// substituted ion's (charge_config, core.py:2553-2557), i.e. V_lat'[s] = V_lat[s] + sum_d dq_d P[s, d].
template <bool COMPACT>
__global__ void __launch_bounds__(256)
vlat_doped_kernel(SysDev S, long long n_traj, int n_dop, const int *__restrict__ dsite,
                  const double *__restrict__ ddq, double *__restrict__ out)
{
    const long long total = n_traj * S.n_sites;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long traj = i / S.n_sites;
        const Site x = make_site<COMPACT>(S, (int)(i - traj * S.n_sites));
        double v = ld_vlat<COMPACT>(S, x);
        for (int d = 0; d < n_dop; ++d) {
            const int site = dsite[traj * n_dop + d];
            if (site >= 0) v = fma(ddq[traj * n_dop + d], ld_pair<COMPACT>(S, x, make_site<COMPACT>(S, site)), v);
        }
        out[i] = v;
    }
}

__global__ void site_pack_kernel(unsigned *out, long long n, int nb, int sy, int sz)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cell = (int)(i / nb), b = (int)(i - (long long)cell * nb);
    const int z = cell % sz, y = (cell / sz) % sy, x = cell / (sz * sy);
    out[i] = (unsigned)b | ((unsigned)x << 8) | ((unsigned)y << 16) | ((unsigned)z << 24);
}

// neigh2[e][s][s2] = neigh[site_centre[neigh[e][s]]][s2] (+ packed form for the compact layout)
__global__ void neigh2_kernel(const int *neigh, const int *site_centre, const unsigned *site_pack,
                              long long n_centres, int nn, int *out, unsigned *out_pack)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_centres * nn * nn) return;
    const int s2 = (int)(i % nn);
    const long long es = i / nn;
    const int mid = neigh[es];
    const int v = neigh[(long long)site_centre[mid] * nn + s2];
    out[i] = v;
    out_pack[i] = site_pack ? site_pack[v] : 0u;
}

__global__ void inv4lambda_kernel(const double *lam, int n, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0 / (4.0 * lam[i]);
}

__global__ void neigh_pack_kernel(const int *neigh, const unsigned *site_pack, long long n, unsigned *out)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = site_pack[neigh[i]];
}

// H entries of the cell-0 basis sites: one thread per (b_a, delta, b_y)
__global__ void stencil_table_kernel(const double *__restrict__ Pu, long long n_sites, int n_basis, int sx, int sy,
                                     int sz, int ncb, int nn, int nnp, const int *__restrict__ cb_all,
                                     const int *__restrict__ nb_all, const int *__restrict__ nb_cell,
                                     double *__restrict__ H)
{
    const int wy = 2 * sy - 1, wz = 2 * sz - 1;
    const long long n_delta = (long long)(2 * sx - 1) * wy * wz;
    const long long total = (long long)ncb * n_delta * ncb;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int by = (int)(i % ncb);
    const long long r = i / ncb;
    const long long dl = r % n_delta;
    const int ba = (int)(r / n_delta);
    const int iz = (int)(dl % wz), iy = (int)((dl / wz) % wy), ix = (int)(dl / ((long long)wz * wy));
    const int cx = (ix + 1) % sx, cy = (iy + 1) % sy, cz = (iz + 1) % sz;   // (d + s) mod s, d = i - (s - 1)
    const int a_all = cb_all[ba], y_all = cb_all[by];
    const double pay = Pu[(long long)a_all * n_sites + ((long long)(cx * sy + cy) * sz + cz) * n_basis + y_all];
    double *out = H + i * nnp;
    for (int d = 0; d < nnp; ++d) {
        double v = 0.0;
        if (d < nn) {
            const int n_all = nb_all[ba * nn + d];
            const int *nc = nb_cell + (ba * nn + d) * 3;
            const int rx = (cx - nc[0] + sx) % sx, ry = (cy - nc[1] + sy) % sy, rz = (cz - nc[2] + sz) % sz;
            const double pny = Pu[(long long)n_all * n_sites + ((long long)(rx * sy + ry) * sz + rz) * n_basis + y_all];
            v = __dsub_rn(pny, pay);
        }
        out[d] = v;
    }
}

// per-(basis, direction) constants: everything of a process except the carrier sum
__global__ void stencil_const_kernel(const double *__restrict__ Pu, long long n_sites, int n_basis, int sy, int sz,
                                     int ncb, int nn, const int *__restrict__ cb_all, const int *__restrict__ nb_all,
                                     const int *__restrict__ nb_cell, const double *__restrict__ v_lat,
                                     const double *__restrict__ e_rel, const int *__restrict__ site_class,
                                     const double *__restrict__ lam, const double *__restrict__ vab,
                                     const double *__restrict__ i4l, double qc, double *__restrict__ cst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncb * nn) return;
    const int b = i / nn, d = i - b * nn;
    const int a_all = cb_all[b], n_all = nb_all[i];
    const int *nc = nb_cell + i * 3;
    const long long n_site = ((long long)(nc[0] * sy + nc[1]) * sz + nc[2]) * n_basis + n_all;
    const double paa = Pu[(long long)a_all * n_sites + a_all];
    const double pab = Pu[(long long)a_all * n_sites + n_site];
    const int cls = site_class[a_all];
    double *o = cst + (long long)b * ST_ROWS * nn + d;
    o[ST_T02 * nn] = __dmul_rn(qc, __dsub_rn(paa, pab));                 // core.py:2010-2014
    o[ST_SHIFT * nn] = __dsub_rn(e_rel[n_site], e_rel[a_all]);           // core.py:2023-2025
    o[ST_LAM * nn] = lam[cls * nn + d];
    o[ST_VAB * nn] = vab[cls * nn + d];
    o[ST_I4L * nn] = i4l[cls * nn + d];
    o[ST_VL * nn] = __dsub_rn(v_lat[n_all], v_lat[a_all]);
}

}  // namespace pycd

#include "kmc_stencil_launch.h"

using namespace pycd;

struct pycd_kmc_system {
    pycd_ctx *ctx = nullptr;
    SysDev dev{};
    long long n_centres = 0;
    int n_class = 0;
    InBuf<double> P;
    InBuf<int> site_centre, site_class, neigh;
    InBuf<double> hopvec, lam, vab, e_rel;
    DevBuf<double> v_lat, i4l;
    DevBuf<unsigned> site_pack, neigh_pack, neigh2_pack;
    DevBuf<int> neigh2;
    bool compact = false;
    std::vector<int> site_centre_h;   // host copy: validates every occupancy array that reaches the device
    // lattice-stencil tables (kmc_stencil.cuh); stencil_why says why they are absent
    StencilDev st{};
    bool stencil_ok = false;
    std::string stencil_why = "not built";
    size_t st_smem = 0;
    DevBuf<double> st_H, st_cst;
    DevBuf<int> st_ctr_key, st_ctr_site, st_nbr_key, st_nbr_ctr, st_cb_all, st_nb_all, st_nb_cell;
    DevBuf<unsigned long long> st_perm, st_nbr_perm;
};

struct pycd_kmc_ensemble {
    pycd_kmc_system *sys = nullptr;
    EnsDev dev{};
    DevBuf<int> occ, done;
    DevBuf<double> t, disp, row, drift, rates, unwrapped, kT_traj, field_traj, dt_grid_traj, energy, energy_grid, dg0_grid;
    DevBuf<double> e_rel_traj, v_lat_traj;   // doped ensembles only
    std::vector<double> energy0;
    DevBuf<long long> start_idx, n_steps, near_tie, clamped;
    std::string last_kernel;
    long long n_active = -1;   // unfinished trajectories after the last advance (-1: all)
    // pipelined read-back (pycd_kmc_read_begin / _end): the grid being copied out while the next batch
    // fills the other one, and snapshots of the small per-trajectory state
    DevBuf<double> unwrapped_alt, snap_t, snap_drift;
    DevBuf<long long> snap_n_steps, snap_near_tie, snap_clamped;
    DevBuf<int> snap_occ;
    cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
    bool read_in_flight = false;
    bool needs_reset = false;   // the displacement grid was handed to a pipelined read: re-arm before advancing
    // pycd_kmc_ensemble_reset without a host synchronisation: the new sites are staged in page-locked memory
    // owned by the ensemble (the caller's buffer is free when the call returns), ev_stage guards its re-use
    int *occ_stage = nullptr;
    cudaEvent_t ev_stage = nullptr;
    bool stage_in_flight = false;
    DevBuf<long long> ones;     // start_path_index = 1 per trajectory (core.py:2781), source of the re-arm copy
    ~pycd_kmc_ensemble() {
        if (ev_snap) cudaEventDestroy(ev_snap);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (ev_stage) cudaEventDestroy(ev_stage);
        if (occ_stage) cudaFreeHost(occ_stage);
    }
};

// occupancy arrays index device tables: every entry must be a site of the carrier's element
static void validate_occupancy(const pycd_kmc_system *sys, const int32_t *occ, size_t count, std::vector<int> &host) {
    host.resize(count);
    PYCD_CUDA(cudaMemcpy(host.data(), occ, sizeof(int) * count,
                         is_device_pointer(occ) ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
    const long long n = sys->dev.n_sites;
    for (int v : host)
        PYCD_REQUIRE(v >= 0 && v < n && sys->site_centre_h[v] >= 0, "occupancy entry is not a carrier site");
}

// energy outputs at t = 0: energy_array[0] = initial energy, everything else 0 (core.py:2715-2718, 2782-2783)
static void arm_energy(pycd_kmc_ensemble *ens, cudaStream_t s);

// Builds the lattice-stencil tables of kmc_stencil.cuh for the unit-rows layout.  Throws Error
// with the reason when the system does not qualify (the caller then keeps the gather kernels):
// the tables need centre = cell*ncb + basis numbering, periodic site classes / site energies, and
// neighbour lists that are translations of those of unit cell 0 up to a permutation inside a
// hop-distance class.
static void setup_stencil(pycd_kmc_system *sys, const pycd_kmc_system_desc *d, cudaStream_t s)
{
    pycd_ctx *ctx = sys->ctx;
    const int nn = d->nn, nb = d->n_basis;
    const long long n = d->n_sites, nc = d->n_centres;
    const int sx = d->size[0], sy = d->size[1], sz = d->size[2];
    const long long cells = (long long)sx * sy * sz;
    auto need = [](bool ok, const char *why) { if (!ok) throw Error(why); };
    need(nn == 4 || nn == 8 || nn == 12, "stencil kernels are instantiated for 4, 8 and 12 neighbour slots");
    need(nc < (1 << 24) && nc % cells == 0, "centre count is not a multiple of the cell count");
    const int ncb = (int)(nc / cells);
    need(ncb >= 1 && ncb <= 127, "more than 127 carrier sites per unit cell");
    const int nnp = (nn + 3) & ~3;
    const int wx = 2 * sx - 1, wy = 2 * sy - 1, wz = 2 * sz - 1;
    const long long n_delta = (long long)wx * wy * wz;
    const long long rs = n_delta * ncb, entries = rs * ncb;
    double max_mb = 96.0;
    if (const char *e = getenv("PYCD_STENCIL_MAX_MB")) max_mb = atof(e);
    need(entries < (1ll << 30) && (double)entries * nnp * 8.0 <= max_mb * 1048576.0,
         "stencil table larger than PYCD_STENCIL_MAX_MB");
    const size_t smem = (size_t)ncb * (ST_ROWS + 3) * nn * sizeof(double);   // raw rows + folded constants
    need(smem <= 32 * 1024, "constant table does not fit in shared memory");

    std::vector<int> neigh((size_t)nc * nn), site_centre((size_t)n), site_class((size_t)n);
    std::vector<double> hop((size_t)nc * nn * 3), e_rel((size_t)n), lam((size_t)d->n_class * nn),
        vab((size_t)d->n_class * nn);
    PYCD_CUDA(cudaStreamSynchronize(s));
    PYCD_CUDA(cudaMemcpy(neigh.data(), sys->neigh.p, sizeof(int) * neigh.size(), cudaMemcpyDeviceToHost));
    PYCD_CUDA(cudaMemcpy(site_centre.data(), sys->site_centre.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    PYCD_CUDA(cudaMemcpy(site_class.data(), sys->site_class.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    PYCD_CUDA(cudaMemcpy(hop.data(), sys->hopvec.p, sizeof(double) * hop.size(), cudaMemcpyDeviceToHost));
    PYCD_CUDA(cudaMemcpy(e_rel.data(), sys->e_rel.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    PYCD_CUDA(cudaMemcpy(lam.data(), sys->lam.p, sizeof(double) * lam.size(), cudaMemcpyDeviceToHost));
    PYCD_CUDA(cudaMemcpy(vab.data(), sys->vab.p, sizeof(double) * vab.size(), cudaMemcpyDeviceToHost));

    std::vector<int> ctr_site((size_t)nc, -1);
    for (long long i = 0; i < n; ++i) {
        const int e = site_centre[i];
        if (e < 0) continue;
        need(e < nc && ctr_site[e] < 0, "site_centre is not a bijection onto the centres");
        ctr_site[e] = (int)i;
    }
    std::vector<int> cb_all(ncb);
    for (int b = 0; b < ncb; ++b) {
        need(ctr_site[b] >= 0 && ctr_site[b] < nb, "centres of unit cell 0 are not numbered first");
        cb_all[b] = ctr_site[b];
    }
    for (long long e = 0; e < nc; ++e)
        need(ctr_site[e] == (e / ncb) * nb + cb_all[e % ncb], "centre numbering is not cell*ncb + basis");
    for (long long i = 0; i < n; ++i)
        need(site_class[i] == site_class[i % nb] && e_rel[i] == e_rel[i % nb],
             "site classes / relative site energies are not lattice periodic");

    auto cell_xyz = [&](long long cell, int *c) {
        c[2] = (int)(cell % sz); c[1] = (int)((cell / sz) % sy); c[0] = (int)(cell / ((long long)sz * sy));
    };
    auto key_of = [&](long long e) {
        int c[3];
        cell_xyz(e / ncb, c);
        return (int)((((long long)c[0] * wy + c[1]) * wz + c[2]) * ncb + e % ncb);
    };
    // canonical directions of a basis site: the slots of its copy in unit cell 0
    std::vector<int> nb_all((size_t)ncb * nn), nb_cell((size_t)ncb * nn * 3);
    for (int b = 0; b < ncb; ++b)
        for (int k = 0; k < nn; ++k) {
            const int ns = neigh[(size_t)b * nn + k];
            need(ns >= 0 && ns < n && site_centre[ns] >= 0, "neighbour is not a carrier site");
            nb_all[b * nn + k] = ns % nb;
            cell_xyz(ns / nb, &nb_cell[(size_t)(b * nn + k) * 3]);
        }
    std::vector<int> ctr_key((size_t)nc), nbr_key((size_t)nc * nn), nbr_ctr((size_t)nc * nn);
    std::vector<unsigned long long> perm((size_t)nc);
    const int size[3] = {sx, sy, sz};
    for (long long e = 0; e < nc; ++e) {
        const int b = (int)(e % ncb);
        int ec[3];
        cell_xyz(e / ncb, ec);
        ctr_key[e] = key_of(e);
        const int cls = site_class[ctr_site[e]];
        need(cls >= 0 && cls < d->n_class, "bad site class");
        unsigned used = 0;
        unsigned long long pm = 0;
        for (int k = 0; k < nn; ++k) {
            const int ns = neigh[(size_t)e * nn + k];
            need(ns >= 0 && ns < n && site_centre[ns] >= 0, "neighbour is not a carrier site");
            int c2[3], rel[3];
            cell_xyz(ns / nb, c2);
            for (int a = 0; a < 3; ++a) rel[a] = ((c2[a] - ec[a]) % size[a] + size[a]) % size[a];
            const double *hv = &hop[((size_t)e * nn + k) * 3];
            int found = -1;
            for (int dd = 0; dd < nn && found < 0; ++dd) {
                if (used & (1u << dd)) continue;
                const int *cc = &nb_cell[(size_t)(b * nn + dd) * 3];
                if (nb_all[b * nn + dd] != ns % nb || cc[0] != rel[0] || cc[1] != rel[1] || cc[2] != rel[2]) continue;
                if (lam[cls * nn + dd] != lam[cls * nn + k] || vab[cls * nn + dd] != vab[cls * nn + k]) continue;
                const double *h0 = &hop[((size_t)b * nn + dd) * 3];
                bool close = true;
                for (int a = 0; a < 3; ++a) close = close && std::fabs(hv[a] - h0[a]) <= 1e-9 * (1.0 + std::fabs(h0[a]));
                if (close) found = dd;
            }
            need(found >= 0, "neighbour lists are not translations of those of unit cell 0");
            used |= 1u << found;
            pm |= (unsigned long long)k << (4 * found);
            const int e2 = site_centre[ns];
            nbr_key[(size_t)e * nn + k] = key_of(e2);
            nbr_ctr[(size_t)e * nn + k] = e2 | ((e2 % ncb) << 24);
        }
        perm[e] = pm;
    }

    auto up_i = [&](DevBuf<int> &b, const std::vector<int> &v) {
        b.alloc(v.size());
        PYCD_CUDA(cudaMemcpyAsync(b.p, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, s));
    };
    up_i(sys->st_ctr_key, ctr_key); up_i(sys->st_ctr_site, ctr_site); up_i(sys->st_nbr_key, nbr_key);
    up_i(sys->st_nbr_ctr, nbr_ctr); up_i(sys->st_cb_all, cb_all); up_i(sys->st_nb_all, nb_all);
    up_i(sys->st_nb_cell, nb_cell);
    sys->st_perm.alloc(perm.size());
    PYCD_CUDA(cudaMemcpyAsync(sys->st_perm.p, perm.data(), sizeof(unsigned long long) * perm.size(), cudaMemcpyHostToDevice, s));
    // the permutation of every neighbour beside its key: a hop finds its new site's permutation in the shared
    // tables instead of waiting for a global load on the owner's chain
    std::vector<unsigned long long> nbr_perm((size_t)nc * nn);
    for (size_t i = 0; i < nbr_perm.size(); ++i) nbr_perm[i] = perm[(size_t)(nbr_ctr[i] & 0xffffff)];
    sys->st_nbr_perm.alloc(nbr_perm.size());
    PYCD_CUDA(cudaMemcpyAsync(sys->st_nbr_perm.p, nbr_perm.data(), sizeof(unsigned long long) * nbr_perm.size(), cudaMemcpyHostToDevice, s));
    sys->st_H.alloc((size_t)entries * nnp);
    sys->st_cst.alloc((size_t)ncb * ST_ROWS * nn);
    stencil_table_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, s>>>(
        sys->P.p, n, nb, sx, sy, sz, ncb, nn, nnp, sys->st_cb_all.p, sys->st_nb_all.p, sys->st_nb_cell.p, sys->st_H.p);
    check_launch(ctx, "stencil_table_kernel");
    stencil_const_kernel<<<(ncb * nn + 127) / 128, 128, 0, s>>>(
        sys->P.p, n, nb, sy, sz, ncb, nn, sys->st_cb_all.p, sys->st_nb_all.p, sys->st_nb_cell.p, sys->v_lat.p,
        sys->e_rel.p, sys->site_class.p, sys->lam.p, sys->vab.p, sys->i4l.p, d->q_carrier,
        sys->st_cst.p);
    check_launch(ctx, "stencil_const_kernel");
    PYCD_CUDA(cudaStreamSynchronize(s));   // the host vectors above are read by the async copies
    StencilDev &t = sys->st;
    t.H = sys->st_H.p; t.ctr_key = sys->st_ctr_key.p; t.ctr_site = sys->st_ctr_site.p;
    t.nbr_key = sys->st_nbr_key.p; t.nbr_ctr = sys->st_nbr_ctr.p; t.perm = sys->st_perm.p; t.nbr_perm = sys->st_nbr_perm.p; t.cst = sys->st_cst.p;
    t.ncb = ncb;
    t.rs_p1 = (int)(rs + 1);
    t.l0_ncb = (int)(((((long long)(sx - 1) * wy + (sy - 1)) * wz + (sz - 1))) * ncb);
    sys->st_smem = smem;
    sys->stencil_ok = true;
    sys->stencil_why.clear();
}

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
        NvtxRange nvtx("pycd.kmc_system_create");
        DeviceGuard g(ctx);
        auto sys = new pycd_kmc_system();
        try {
            sys->ctx = ctx;
            const size_t n = (size_t)d->n_sites;
            cudaStream_t s = ctx->stream;
            sys->compact = (d->p_layout == PYCD_P_UNIT_ROWS);
            size_t p_rows = n;
            if (sys->compact) {
                const long long cells = (long long)d->size[0] * d->size[1] * d->size[2];
                PYCD_REQUIRE(d->n_basis > 0 && d->n_basis <= 255, "compact layout needs 1..255 sites per unit cell");
                PYCD_REQUIRE(d->size[0] > 0 && d->size[0] <= 255 && d->size[1] > 0 && d->size[1] <= 255 &&
                                 d->size[2] > 0 && d->size[2] <= 255, "compact layout needs 1..255 cells per axis");
                PYCD_REQUIRE(cells * d->n_basis == d->n_sites, "n_basis * cells != n_sites");
                p_rows = (size_t)d->n_basis;
                sys->site_pack.alloc(n);
                site_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sys->site_pack.p, (long long)n,
                                                                             d->n_basis, d->size[1], d->size[2]);
                check_launch(ctx, "site_pack_kernel");
            } else {
                PYCD_REQUIRE(d->p_layout == PYCD_P_DENSE, "unknown p_layout");
            }
            sys->P.bind(d->P, p_rows * n, s);
            sys->site_centre.bind(d->site_centre, n, s);
            sys->site_class.bind(d->site_class, n, s);
            sys->neigh.bind(d->neigh, (size_t)d->n_centres * d->nn, s);
            sys->hopvec.bind(d->hopvec, (size_t)d->n_centres * d->nn * 3, s);
            sys->lam.bind(d->lam, (size_t)d->n_class * d->nn, s);
            sys->vab.bind(d->vab, (size_t)d->n_class * d->nn, s);
            sys->e_rel.bind(d->e_rel, n, s);
            {   // the neighbour table is dereferenced by setup kernels and step kernels: range-check it
                const size_t nn_tot = (size_t)d->n_centres * d->nn;
                std::vector<int> neigh_h(nn_tot);
                sys->site_centre_h.resize(n);
                PYCD_CUDA(cudaStreamSynchronize(s));
                PYCD_CUDA(cudaMemcpy(neigh_h.data(), sys->neigh.p, sizeof(int) * nn_tot, cudaMemcpyDeviceToHost));
                PYCD_CUDA(cudaMemcpy(sys->site_centre_h.data(), sys->site_centre.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
                for (int v : sys->site_centre_h) PYCD_REQUIRE(v >= -1 && v < d->n_centres, "site_centre entry out of range");
                for (int v : neigh_h)
                    PYCD_REQUIRE(v >= 0 && (size_t)v < n && sys->site_centre_h[v] >= 0,
                                 "neigh entry is not a site of the carrier's element");
            }
            if (sys->compact) {
                const long long nn_tot = (long long)d->n_centres * d->nn;
                sys->neigh_pack.alloc((size_t)nn_tot);
                neigh_pack_kernel<<<(unsigned)((nn_tot + 255) / 256), 256, 0, s>>>(sys->neigh.p, sys->site_pack.p,
                                                                                  nn_tot, sys->neigh_pack.p);
                check_launch(ctx, "neigh_pack_kernel");
            }
            {
                const int nl = d->n_class * d->nn;
                sys->i4l.alloc((size_t)nl);
                inv4lambda_kernel<<<(nl + 127) / 128, 128, 0, s>>>(sys->lam.p, nl, sys->i4l.p);
                check_launch(ctx, "inv4lambda_kernel");
            }
            {
                const long long n2 = (long long)d->n_centres * d->nn * d->nn;
                sys->neigh2.alloc((size_t)n2);
                sys->neigh2_pack.alloc((size_t)n2);
                neigh2_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(sys->neigh.p, sys->site_centre.p,
                                                                          sys->site_pack.p, d->n_centres, d->nn,
                                                                          sys->neigh2.p, sys->neigh2_pack.p);
                check_launch(ctx, "neigh2_kernel");
            }
            InBuf<double> q;
            q.bind(d->q_lat, n, s);
            // V_lat = P . q_lat: one entry per site (dense) or per basis site (compact: the
            // lattice charges are periodic, so V_lat[i] = V_lat[basis_i])
            sys->v_lat.alloc(p_rows);
            KernelTimer tv(ctx, KC_VLAT);
            vlat_kernel<<<(unsigned)((p_rows * 32 + 255) / 256), 256, 0, s>>>(sys->P.p, q.p, (long long)p_rows,
                                                                              (long long)n, sys->v_lat.p);
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
            v.site_pack = sys->site_pack.p;
            v.neigh_pack = sys->neigh_pack.p;
            v.i4l = sys->i4l.p;
            v.neigh2 = sys->neigh2.p;
            v.neigh2_pack = sys->neigh2_pack.p;
            v.n_basis = d->n_basis; v.sx = d->size[0]; v.sy = d->size[1]; v.sz = d->size[2];
            sys->n_centres = d->n_centres;
            sys->n_class = d->n_class;
            if (sys->compact) {
                const char *off = getenv("PYCD_STENCIL");
                if (off && std::string(off) == "0") {
                    sys->stencil_why = "disabled by PYCD_STENCIL=0";
                } else {
                    try {
                        setup_stencil(sys, d, s);
                    } catch (const Error &e) {
                        sys->stencil_ok = false;
                        sys->stencil_why = e.what();
                    }
                }
            } else {
                sys->stencil_why = "dense layout";
            }
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
        PYCD_CUDA(cudaMemcpy(v_lat, sys->v_lat.p, sizeof(double) * sys->v_lat.n,
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
        PYCD_REQUIRE(!d->dopant_site || (d->n_dopant_max > 0 && d->dopant_dq), "dopant_site needs n_dopant_max and dopant_dq");
        pycd_ctx *ctx = sys->ctx;
        DeviceGuard g(ctx);
        const long long nt = d->n_traj;
        const int C = d->n_carriers;
        const long long n_proc_ll = (long long)C * sys->dev.nn;
        PYCD_REQUIRE(n_proc_ll <= 2560, "more than 2560 processes per trajectory is not supported");
        // validate initial sites on the host copy (they index device tables)
        std::vector<int> occ_h;
        validate_occupancy(sys, d->occupancy0, (size_t)nt * C, occ_h);
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
            ens->ones.alloc(nt);
            PYCD_CUDA(cudaMemcpyAsync(ens->ones.p, ones.data(), sizeof(long long) * nt, cudaMemcpyHostToDevice, s));
            if (d->record_unwrapped) {
                ens->unwrapped.alloc((size_t)nt * d->n_path * 3 * C);
                ens->unwrapped.zero(s);
            }
            if (d->energy0) {
                ens->energy0.resize(nt);
                PYCD_CUDA(cudaMemcpy(ens->energy0.data(), d->energy0, sizeof(double) * nt, cudaMemcpyDefault));
                ens->energy.alloc(nt);
                ens->energy_grid.alloc((size_t)nt * d->n_path);
                ens->dg0_grid.alloc((size_t)nt * d->n_path);
            }
            if (d->kT_traj) {
                ens->kT_traj.alloc(nt);
                PYCD_CUDA(cudaMemcpyAsync(ens->kT_traj.p, d->kT_traj, sizeof(double) * nt, cudaMemcpyDefault, s));
            }
            if (d->dt_grid_traj) {
                std::vector<double> dtg((size_t)nt);
                PYCD_CUDA(cudaMemcpy(dtg.data(), d->dt_grid_traj, sizeof(double) * nt, cudaMemcpyDefault));
                for (double v : dtg) PYCD_REQUIRE(v > 0, "bad per-trajectory time grid");
                ens->dt_grid_traj.alloc(nt);
                PYCD_CUDA(cudaMemcpyAsync(ens->dt_grid_traj.p, dtg.data(), sizeof(double) * nt, cudaMemcpyHostToDevice, s));
                PYCD_CUDA(cudaStreamSynchronize(s));
            }
            if (d->field_traj) {
                ens->field_traj.alloc((size_t)nt * 3);
                PYCD_CUDA(cudaMemcpyAsync(ens->field_traj.p, d->field_traj, sizeof(double) * nt * 3, cudaMemcpyDefault, s));
            }
            if (d->e_rel_traj || d->dopant_site) {   // doping hooks
                const long long n = sys->dev.n_sites;
                ens->e_rel_traj.alloc((size_t)nt * n);
                if (d->e_rel_traj)
                    PYCD_CUDA(cudaMemcpyAsync(ens->e_rel_traj.p, d->e_rel_traj, sizeof(double) * nt * n, cudaMemcpyDefault, s));
                else
                    for (long long i = 0; i < nt; ++i)
                        PYCD_CUDA(cudaMemcpyAsync(ens->e_rel_traj.p + i * n, sys->dev.e_rel, sizeof(double) * n,
                                                  cudaMemcpyDeviceToDevice, s));
                const int nd = d->dopant_site ? d->n_dopant_max : 0;
                DevBuf<int> dsite;
                DevBuf<double> ddq;
                if (nd > 0) {
                    std::vector<int> site_h((size_t)nt * nd);
                    PYCD_CUDA(cudaMemcpy(site_h.data(), d->dopant_site, sizeof(int) * nt * nd, cudaMemcpyDefault));
                    for (int v : site_h) PYCD_REQUIRE(v >= -1 && v < n, "dopant site out of range");
                    dsite.alloc((size_t)nt * nd);
                    ddq.alloc((size_t)nt * nd);
                    PYCD_CUDA(cudaMemcpyAsync(dsite.p, site_h.data(), sizeof(int) * nt * nd, cudaMemcpyHostToDevice, s));
                    PYCD_CUDA(cudaMemcpyAsync(ddq.p, d->dopant_dq, sizeof(double) * nt * nd, cudaMemcpyDefault, s));
                }
                ens->v_lat_traj.alloc((size_t)nt * n);
                const unsigned grid = (unsigned)std::min<long long>((nt * n + 255) / 256, 148ll * 16);
                if (sys->compact) vlat_doped_kernel<true><<<grid, 256, 0, s>>>(sys->dev, nt, nd, dsite.p, ddq.p, ens->v_lat_traj.p);
                else vlat_doped_kernel<false><<<grid, 256, 0, s>>>(sys->dev, nt, nd, dsite.p, ddq.p, ens->v_lat_traj.p);
                check_launch(ctx, "vlat_doped_kernel");
                PYCD_CUDA(cudaStreamSynchronize(s));   // dsite / ddq are released at the end of this block
            }
            EnsDev &e = ens->dev;
            e.e_rel_traj = ens->e_rel_traj.p; e.v_lat_traj = ens->v_lat_traj.p;
            e.energy = ens->energy.p; e.energy_grid = ens->energy_grid.p; e.dg0_grid = ens->dg0_grid.p;
            e.n_path = d->n_path;
            arm_energy(ens, s);
            PYCD_CUDA(cudaStreamSynchronize(s));
            e.C = C; e.n_proc = (int)n_proc_ll; e.n_traj = nt; e.traj_id0 = d->traj_id0;
            e.occ = ens->occ.p; e.t = ens->t.p; e.start_idx = ens->start_idx.p; e.done = ens->done.p;
            e.disp = ens->disp.p; e.row = ens->row.p; e.n_steps = ens->n_steps.p;
            e.near_tie = ens->near_tie.p; e.clamped = ens->clamped.p; e.drift = ens->drift.p;
            e.rates = ens->rates.p; e.kT_traj = ens->kT_traj.p; e.field_traj = ens->field_traj.p; e.dt_grid_traj = ens->dt_grid_traj.p;
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

static void arm_energy(pycd_kmc_ensemble *ens, cudaStream_t s) {
    if (!ens->energy.p) return;
    const size_t nt = ens->energy0.size();
    const long long n_path = ens->dev.n_path;
    ens->energy_grid.zero(s);
    ens->dg0_grid.zero(s);
    PYCD_CUDA(cudaMemcpyAsync(ens->energy.p, ens->energy0.data(), sizeof(double) * nt, cudaMemcpyHostToDevice, s));
    PYCD_CUDA(cudaMemcpy2DAsync(ens->energy_grid.p, sizeof(double) * n_path, ens->energy0.data(), sizeof(double),
                                sizeof(double), nt, cudaMemcpyHostToDevice, s));
}

extern "C" int pycd_kmc_read_energy(pycd_kmc_ensemble *ens, double *energy_grid, double *dg0_grid) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        PYCD_REQUIRE(ens->energy.p, "ensemble was created without energy0");
        DeviceGuard g(ens->sys->ctx);
        cudaStream_t s = ens->sys->ctx->stream;
        const size_t n = (size_t)ens->dev.n_traj * ens->dev.n_path;
        if (energy_grid) PYCD_CUDA(cudaMemcpyAsync(energy_grid, ens->energy_grid.p, sizeof(double) * n, cudaMemcpyDefault, s));
        if (dg0_grid) PYCD_CUDA(cudaMemcpyAsync(dg0_grid, ens->dg0_grid.p, sizeof(double) * n, cudaMemcpyDefault, s));
        PYCD_CUDA(cudaStreamSynchronize(s));
    });
}

extern "C" int pycd_kmc_ensemble_reset(pycd_kmc_ensemble *ens, const int32_t *occupancy0, uint64_t traj_id0) {
    return guarded([&] {
        PYCD_REQUIRE(ens && occupancy0, "NULL argument");
        pycd_ctx *ctx = ens->sys->ctx;
        DeviceGuard g(ctx);
        EnsDev &E = ens->dev;
        const size_t nt = (size_t)E.n_traj;
        cudaStream_t s = ctx->stream;
        // stream-ordered, no host synchronisation: the re-arm of batch k+1 can be enqueued while batch k runs
        if (!ens->occ_stage) {
            PYCD_CUDA(cudaHostAlloc((void **)&ens->occ_stage, sizeof(int) * nt * E.C, cudaHostAllocDefault));
            PYCD_CUDA(cudaEventCreateWithFlags(&ens->ev_stage, cudaEventDisableTiming));
        }
        if (ens->stage_in_flight) PYCD_CUDA(cudaEventSynchronize(ens->ev_stage));   // previous re-arm has left the stage
        std::vector<int> occ_h;
        validate_occupancy(ens->sys, occupancy0, nt * E.C, occ_h);
        memcpy(ens->occ_stage, occ_h.data(), sizeof(int) * nt * E.C);
        PYCD_CUDA(cudaMemcpyAsync(ens->occ.p, ens->occ_stage, sizeof(int) * nt * E.C, cudaMemcpyHostToDevice, s));
        PYCD_CUDA(cudaEventRecord(ens->ev_stage, s));
        ens->stage_in_flight = true;
        ens->done.zero(s); ens->t.zero(s); ens->disp.zero(s); ens->row.zero(s); ens->drift.zero(s);
        ens->rates.zero(s); ens->n_steps.zero(s); ens->near_tie.zero(s); ens->clamped.zero(s);
        ens->unwrapped.zero(s);
        PYCD_CUDA(cudaMemcpyAsync(ens->start_idx.p, ens->ones.p, sizeof(long long) * nt, cudaMemcpyDeviceToDevice, s));
        E.traj_id0 = traj_id0;
        ens->n_active = -1;
        ens->needs_reset = false;
        arm_energy(ens, s);
    });
}

extern "C" int pycd_kmc_ensemble_destroy(pycd_kmc_ensemble *ens) {
    return guarded([&] {
        if (!ens) return;
        DeviceGuard g(ens->sys->ctx);
        if (ens->read_in_flight) cudaEventSynchronize(ens->ev_copied);
        delete ens;
    });
}

template <int BS, bool COMPACT, int CT, int NNT>
static void launch_step_impl(pycd_ctx *ctx, const SysDev &S, const EnsDev &E, const AdvanceArgs &A, size_t smem) {
    auto kern = kmc_step_kernel<BS, COMPACT, CT, NNT>;
    if (smem > 48 * 1024)
        PYCD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)E.n_traj, BS, smem, ctx->stream>>>(S, E, A);
    check_launch(ctx, "kmc_step_kernel");
}

template <int BS>
static void launch_step(pycd_ctx *ctx, bool compact, const SysDev &S, const EnsDev &E, const AdvanceArgs &A,
                        size_t smem) {
    if (compact) launch_step_impl<BS, true, 0, 0>(ctx, S, E, A, smem);
    else launch_step_impl<BS, false, 0, 0>(ctx, S, E, A, smem);
}

// body of pycd_kmc_advance / pycd_kmc_advance_async.  in_flight = true: no host buffers, no host
// synchronisation; the kernel timer is left with the context and the finished flags are read by pycd_kmc_wait
static int advance_impl(pycd_kmc_ensemble *ens, int64_t max_steps, const double *draws,
                        int32_t *events_out, double *times_out, int64_t *steps_done,
                        int64_t *n_active, bool in_flight) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        PYCD_REQUIRE(max_steps > 0, "max_steps must be positive");
        PYCD_REQUIRE(!ens->needs_reset, "the ensemble's grid was handed to pycd_kmc_read_begin: call "
                                        "pycd_kmc_ensemble_reset before advancing again");
        NvtxRange nvtx("pycd.kmc_advance");
        pycd_ctx *ctx = ens->sys->ctx;
        DeviceGuard g(ctx);
        EnsDev &E = ens->dev;
        PYCD_REQUIRE(E.refresh_interval <= 1 || max_steps % E.refresh_interval == 0,
                     "max_steps must be a multiple of refresh_interval");
        PYCD_REQUIRE(E.rng_mode != PYCD_RNG_REPLAY || draws, "REPLAY mode needs draws");
        PYCD_REQUIRE(!in_flight || E.rng_mode != PYCD_RNG_REPLAY, "pycd_kmc_advance_async needs a counter-based RNG mode");
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
        const size_t smem = kmc_smem_bytes(E.n_proc, E.C, ens->sys->dev.nn);
        PYCD_REQUIRE(smem <= 200 * 1024, "trajectory state does not fit in shared memory");
        const bool cp = ens->sys->compact;
        KernelTimer tk(ctx, KC_KMC_STEP, in_flight);
        // PYCD_KMC_VARIANT=process keeps the one-thread-per-process kernel for A/B measurements
        const char *kv = getenv("PYCD_KMC_VARIANT");
        const bool per_process = kv && std::string(kv) == "process";
        const std::string variant = kv ? kv : "";
        // the lattice-stencil kernels cover full-PBC systems with 4 / 8 / 12 neighbour slots; doped
        // ensembles (per-trajectory site energies / lattice potential) run on their featured variant
        const bool doped = E.v_lat_traj != nullptr;
        const int nn_sys = ens->sys->dev.nn;
        const int c_max = (nn_sys == 4) ? 128 : 64;
        const bool use_stencil = cp && ens->sys->stencil_ok && E.C <= c_max &&
                                 (variant.empty() || variant == "stencil" || variant == "stencil_1warp");
        if (variant == "stencil" && !use_stencil)
            throw Error("PYCD_KMC_VARIANT=stencil: stencil kernel unavailable (" +
                        (ens->sys->stencil_ok ? std::string("shape not covered") : ens->sys->stencil_why) + ")");
        ens->last_kernel = "kmc_step_kernel";
        int bs_force = 0;   // diagnostic: PYCD_KMC_BS forces the block size of the generic kernel
        if (const char *e = getenv("PYCD_KMC_BS")) bs_force = atoi(e);
        if (use_stencil && bs_force == 0) {
            // one CTA of one or two warps per trajectory over the lattice-stencil table (kmc_stencil.cuh)
            const unsigned g = (unsigned)E.n_traj;
            const size_t sm = ens->sys->st_smem;
            // small ensembles (a few trajectories per SM) are bound by the latency of a step: two warps
            // per trajectory; large ones by instructions per step: one warp, two carriers per lane
            // (finished trajectories leave their CTA at once, so the ACTIVE count decides)
            const long long live = ens->n_active >= 0 ? ens->n_active : (long long)E.n_traj;
            int nwc, cpl;
            if (E.C <= 32) { nwc = 1; cpl = 1; }
            else if (E.C <= 64) {
                const bool wide = nn_sys != 4 || (live <= 4ll * ctx->n_sm && variant != "stencil_1warp");
                nwc = wide ? 2 : 1;
                cpl = wide ? 1 : 2;
            } else { nwc = 2; cpl = 2; }
            bool ok = false;
            if (nn_sys == 4) ok = stencil_launch_nn4(ctx, nwc, cpl, g, sm, ens->sys->dev, ens->sys->st, E, A);
            else if (nn_sys == 8) ok = stencil_launch_nn8(ctx, nwc, cpl, g, sm, ens->sys->dev, ens->sys->st, E, A);
            else if (nn_sys == 12) ok = stencil_launch_nn12(ctx, nwc, cpl, g, sm, ens->sys->dev, ens->sys->st, E, A);
            if (!ok) throw Error("stencil kernel shape not instantiated");
            check_launch(ctx, "kmc_step_warp_kernel");
            ens->last_kernel = "kmc_step_warp_kernel<" + std::to_string(nwc) + "," + std::to_string(cpl) + "," +
                               std::to_string(nn_sys) + ">";
        }
        else if (bs_force == 32) launch_step<32>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (bs_force == 64) launch_step<64>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (bs_force == 128) launch_step<128>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (bs_force == 256) launch_step<256>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (E.n_proc <= 32) launch_step<32>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (E.n_proc <= 64) launch_step<64>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (E.n_proc <= 128) launch_step<128>(ctx, cp, ens->sys->dev, E, A, smem);
        else if (E.C == 64 && ens->sys->dev.nn == 4 && !per_process && !doped) {
            // the benchmark shape: one thread per carrier, process state in registers
            // between 4 and 8 CTAs per SM the 128-register variant keeps every trajectory resident in
            // one wave (measured +12 % at 1024 trajectories); beyond that both run in waves
            const bool lean = E.n_traj > 4ll * ctx->n_sm && E.n_traj <= 8ll * ctx->n_sm;
            const unsigned g = (unsigned)E.n_traj;
            if (cp && lean) kmc_step_carrier_kernel<64, 4, true, true><<<g, 64, 0, ctx->stream>>>(ens->sys->dev, E, A);
            else if (cp) kmc_step_carrier_kernel<64, 4, true, false><<<g, 64, 0, ctx->stream>>>(ens->sys->dev, E, A);
            else if (lean) kmc_step_carrier_kernel<64, 4, false, true><<<g, 64, 0, ctx->stream>>>(ens->sys->dev, E, A);
            else kmc_step_carrier_kernel<64, 4, false, false><<<g, 64, 0, ctx->stream>>>(ens->sys->dev, E, A);
            check_launch(ctx, "kmc_step_carrier_kernel");
            ens->last_kernel = lean ? "kmc_step_carrier_kernel<64,4,lean>" : "kmc_step_carrier_kernel<64,4>";
        } else if (E.C == 64 && ens->sys->dev.nn == 4 && cp)   // one thread per process, fully specialised
            launch_step_impl<256, true, 64, 4>(ctx, ens->sys->dev, E, A, smem);
        else launch_step<256>(ctx, cp, ens->sys->dev, E, A, smem);
        tk.stop(1);
        if (in_flight) {
            tk.defer();
            return;
        }
        ev.finish(s);
        tm.finish(s);
        sdn.finish(s);
        std::vector<int> done_h(nt);
        PYCD_CUDA(cudaMemcpyAsync(done_h.data(), ens->done.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, s));
        PYCD_CUDA(cudaStreamSynchronize(s));
        tk.read();
        {
            int64_t act = 0;
            for (int v : done_h) act += (v == 0);
            ens->n_active = act;
            if (n_active) *n_active = act;
        }
    });
}

extern "C" int pycd_kmc_advance(pycd_kmc_ensemble *ens, int64_t max_steps, const double *draws,
                                int32_t *events_out, double *times_out, int64_t *steps_done,
                                int64_t *n_active) {
    return advance_impl(ens, max_steps, draws, events_out, times_out, steps_done, n_active, false);
}

extern "C" int pycd_kmc_advance_async(pycd_kmc_ensemble *ens, int64_t max_steps) {
    return advance_impl(ens, max_steps, nullptr, nullptr, nullptr, nullptr, nullptr, true);
}

extern "C" int pycd_kmc_wait(pycd_kmc_ensemble *ens, int64_t *n_active) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        pycd_ctx *ctx = ens->sys->ctx;
        DeviceGuard g(ctx);
        const size_t nt = (size_t)ens->dev.n_traj;
        std::vector<int> done_h(nt);
        PYCD_CUDA(cudaMemcpyAsync(done_h.data(), ens->done.p, sizeof(int) * nt, cudaMemcpyDeviceToHost, ctx->stream));
        PYCD_CUDA(cudaStreamSynchronize(ctx->stream));
        collect_timers(ctx);
        int64_t act = 0;
        for (int v : done_h) act += (v == 0);
        ens->n_active = act;
        if (n_active) *n_active = act;
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


// Pipelined read-back.  The state after the launches issued so far is snapshotted on the compute stream;
// the ensemble switches to its second displacement grid, and the copy stream moves the snapshot to the
// HOST buffers while the caller re-arms the ensemble and runs the next batch.
extern "C" int pycd_kmc_read_begin(pycd_kmc_ensemble *ens, double *unwrapped, int64_t *n_steps, double *sim_time,
                                   int32_t *occupancy, double *drift, int64_t *near_tie, int64_t *clamped) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        PYCD_REQUIRE(!ens->read_in_flight, "a pipelined read is in flight: call pycd_kmc_read_end first");
        PYCD_REQUIRE(!ens->energy.p, "pipelined read-back is not available with energy outputs");
        pycd_ctx *ctx = ens->sys->ctx;
        DeviceGuard g(ctx);
        EnsDev &E = ens->dev;
        const size_t nt = (size_t)E.n_traj, C = (size_t)E.C;
        cudaStream_t s = ctx->stream, cs = ctx->copy_stream;
        if (!ens->ev_snap) {
            PYCD_CUDA(cudaEventCreateWithFlags(&ens->ev_snap, cudaEventDisableTiming));
            PYCD_CUDA(cudaEventCreateWithFlags(&ens->ev_copied, cudaEventDisableTiming));
            ens->snap_t.alloc(nt); ens->snap_drift.alloc(nt * 3 * C); ens->snap_n_steps.alloc(nt);
            ens->snap_near_tie.alloc(nt); ens->snap_clamped.alloc(nt); ens->snap_occ.alloc(nt * C);
        }
        const double *grid = nullptr;
        if (unwrapped) {
            PYCD_REQUIRE(ens->unwrapped.p, "ensemble was created without record_unwrapped");
            if (!ens->unwrapped_alt.p) {
                ens->unwrapped_alt.alloc(ens->unwrapped.n);
                ens->unwrapped_alt.zero(s);
            }
            grid = ens->unwrapped.p;
            std::swap(ens->unwrapped.p, ens->unwrapped_alt.p);   // the next batch fills the other grid
            E.unwrapped = ens->unwrapped.p;
            ens->needs_reset = true;
        }
        auto snap = [&](void *dst, const void *src, size_t bytes) {
            PYCD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
        };
        snap(ens->snap_t.p, ens->t.p, sizeof(double) * nt);
        snap(ens->snap_drift.p, ens->drift.p, sizeof(double) * nt * 3 * C);
        snap(ens->snap_n_steps.p, ens->n_steps.p, sizeof(long long) * nt);
        snap(ens->snap_near_tie.p, ens->near_tie.p, sizeof(long long) * nt);
        snap(ens->snap_clamped.p, ens->clamped.p, sizeof(long long) * nt);
        snap(ens->snap_occ.p, ens->occ.p, sizeof(int) * nt * C);
        PYCD_CUDA(cudaEventRecord(ens->ev_snap, s));
        PYCD_CUDA(cudaStreamWaitEvent(cs, ens->ev_snap, 0));
        auto pull = [&](void *dst, const void *src, size_t bytes) {
            if (dst) PYCD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, cs));
        };
        pull(unwrapped, grid, sizeof(double) * nt * E.n_path * 3 * C);
        pull(n_steps, ens->snap_n_steps.p, sizeof(long long) * nt);
        pull(sim_time, ens->snap_t.p, sizeof(double) * nt);
        pull(occupancy, ens->snap_occ.p, sizeof(int) * nt * C);
        pull(drift, ens->snap_drift.p, sizeof(double) * nt * 3 * C);
        pull(near_tie, ens->snap_near_tie.p, sizeof(long long) * nt);
        pull(clamped, ens->snap_clamped.p, sizeof(long long) * nt);
        PYCD_CUDA(cudaEventRecord(ens->ev_copied, cs));
        ens->read_in_flight = true;
    });
}

extern "C" int pycd_kmc_read_end(pycd_kmc_ensemble *ens) {
    return guarded([&] {
        PYCD_REQUIRE(ens, "NULL ensemble");
        if (!ens->read_in_flight) return;
        DeviceGuard g(ens->sys->ctx);
        PYCD_CUDA(cudaEventSynchronize(ens->ev_copied));
        ens->read_in_flight = false;
    });
}

extern "C" int pycd_kmc_last_kernel(pycd_kmc_ensemble *ens, char *buf, int32_t n) {
    return guarded([&] {
        PYCD_REQUIRE(ens && buf && n > 0, "NULL argument");
        snprintf(buf, (size_t)n, "%s", ens->last_kernel.c_str());
    });
}

extern "C" int pycd_kmc_system_stencil(pycd_kmc_system *sys, int32_t *available, int64_t *table_bytes, char *why,
                                       int32_t n) {
    return guarded([&] {
        PYCD_REQUIRE(sys, "NULL system");
        if (available) *available = sys->stencil_ok ? 1 : 0;
        if (table_bytes) *table_bytes = (int64_t)(sys->st_H.n * sizeof(double));
        if (why && n > 0) snprintf(why, (size_t)n, "%s", sys->stencil_why.c_str());
    });
}

extern "C" int pycd_kmc_unwrapped_device(pycd_kmc_ensemble *ens, const double **dev_ptr) {
    return guarded([&] {
        PYCD_REQUIRE(ens && dev_ptr, "NULL argument");
        PYCD_REQUIRE(ens->unwrapped.p, "ensemble was created without record_unwrapped");
        *dev_ptr = ens->unwrapped.p;
    });
}
