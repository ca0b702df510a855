// kmc_types.cuh -- device-side views of a KMC system / ensemble shared by the step-kernel translation
// units (kmc.cu: table setup, gather kernels, C ABI; kmc_stencil_nn*.cu: lattice-stencil step kernels).
#pragma once

#include "common.cuh"

namespace pycd {

struct SysDev {
    const double *P;          // dense: [N][N]; compact: [n_basis][N] (rows of unit cell 0)
    long long n_sites;
    const int *site_centre;
    const int *site_class;
    const unsigned *site_pack;  // compact only: basis | x<<8 | y<<16 | z<<24
    int nn;
    const int *neigh;
    const unsigned *neigh_pack; // compact only: site_pack of neigh[e][slot]
    const int *neigh2;          // [n_centres][nn][nn]: neighbours of the neighbours (carrier kernel)
    const unsigned *neigh2_pack;
    const double *hopvec;
    const double *lam;
    const double *vab;
    const double *i4l;        // 1/(4 lambda) per (class, slot)
    const double *e_rel;
    const double *v_lat;      // dense: [N]; compact: [n_basis] (doped trajectory: [N] either way)
    int vlat_by_site;         // 1: v_lat is indexed by site in the compact layout too (doped trajectory)
    double qc, kT, vn;
    double field[3];
    int field_active;
    int n_basis, sx, sy, sz;  // compact only
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
    const double *dt_grid_traj; // NULL or [n_traj]: per-trajectory time_interval (sweeps: one grid per condition)
    const double *e_rel_traj;  // NULL or [n_traj][N]: a doped trajectory's site energies (core.py:2750-2764)
    const double *v_lat_traj;  // NULL or [n_traj][N]: its lattice potential, dopant charges included
    double *unwrapped;     // [n_traj][n_path][3C] or NULL
    double *energy;        // [n_traj] current_state_energy, or NULL (energy outputs off)
    double *energy_grid;   // [n_traj][n_path]
    double *dg0_grid;      // [n_traj][n_path]
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

// Philox4x32-10; identical to the CPU checker's generator (tests pin both to the
// Random123 known-answer vector)
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


// rows of the per-(basis, direction) constant table
constexpr int ST_T02 = 0, ST_SHIFT = 1, ST_LAM = 2, ST_VAB = 3, ST_I4L = 4, ST_VL = 5, ST_ROWS = 6;

struct StencilDev {
    const double *H;         // [ncb][n_delta][ncb][NNP]
    const int *ctr_key;      // [n_centres] K of a centre
    const int *ctr_site;     // [n_centres] site index of a centre
    const int *nbr_key;      // [n_centres][nn] K of the neighbour in REFERENCE slot s
    const int *nbr_ctr;      // [n_centres][nn] its centre index | basis << 24
    const unsigned long long *perm;   // [n_centres] 4 bits per canonical direction d: reference slot of d
    const unsigned long long *nbr_perm;   // [n_centres][nn] perm of the neighbour in reference slot s (prefetched with the row)
    const double *cst;       // [ncb][ST_ROWS][nn]
    int ncb, rs_p1, l0_ncb;
};

}  // namespace pycd
