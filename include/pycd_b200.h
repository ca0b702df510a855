/*
 * pycd_b200.h -- C ABI of the B200-native PyCD hot path (libpycd_b200.so).
 *
 * The reference (vpasumarthi/PyCD) is pure Python and has no FFI of its own; the
 * drop-in boundary is its driver functions (PyCD/material_setup.py:13,
 * PyCD/material_run.py:13, PyCD/material_msd.py:13) and the files they exchange.
 * The entry points below are what those drivers' bodies bind instead of the
 * reference's numpy loops; each one cites the reference code it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch/numpy types.  Every bulk pointer may be a
 *    HOST pointer (pageable or pinned) or a DEVICE pointer on the context's GPU; the
 *    library detects which (cudaPointerGetAttributes) and stages copies itself.
 *    Host buffers give the end-to-end path, device buffers the resident path.
 *  - arrays are C-contiguous, caller-owned, never retained unless stated.
 *  - every call is blocking: the context's stream is synchronised before return.
 *  - return value 0 = ok, non-zero = error; pycd_last_error() gives the message
 *    (thread-local).  There is NO CPU fallback: pycd_ctx_create fails without an
 *    sm_100 device.
 *  - units are the reference's atomic units (bohr, Hartree, a.u. time).
 */
#ifndef PYCD_B200_H
#define PYCD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYCD_ABI_VERSION 3

typedef struct pycd_ctx pycd_ctx;
typedef struct pycd_kmc_system pycd_kmc_system;
typedef struct pycd_kmc_ensemble pycd_kmc_ensemble;

/* ---- context ----------------------------------------------------------- */
int pycd_abi_version(void);
const char *pycd_last_error(void);
/* device = CUDA ordinal (one process per GPU: pass LOCAL_RANK). */
int pycd_ctx_create(int device, pycd_ctx **out);
int pycd_ctx_destroy(pycd_ctx *ctx);
/* number of SMs, free/total bytes of the context's device */
int pycd_ctx_info(pycd_ctx *ctx, int32_t *n_sm, int64_t *free_bytes, int64_t *total_bytes);
/* total kernel launches issued by this context so far (bench "gpu_launches") */
int64_t pycd_ctx_launch_count(pycd_ctx *ctx);
/* elapsed device milliseconds (CUDA events on the context's stream) of the kernels
 * of the most recent compute call, by kernel class: 0 = ewald fourier, 1 = ewald
 * finish, 2 = ewald expand, 3 = kmc step kernel, 4 = msd, 5 = v_lat matvec */
double pycd_ctx_last_kernel_ms(pycd_ctx *ctx, int32_t kernel_class);
/* accumulated device milliseconds / launches of a kernel class since the last reset */
double pycd_ctx_total_kernel_ms(pycd_ctx *ctx, int32_t kernel_class);
int64_t pycd_ctx_class_launches(pycd_ctx *ctx, int32_t kernel_class);
int pycd_ctx_reset_timers(pycd_ctx *ctx);
/* evict the L2 (512 MB memset enqueued on the context's stream, no host synchronisation); for cold-cache timing */
int pycd_ctx_flush_l2(pycd_ctx *ctx);

/* Page-locked host buffers for the end-to-end path (host numpy views over them make the
 * H2D/D2H copies of a call run at PCIe speed instead of through pageable staging). */
int pycd_host_alloc(int64_t bytes, void **out);
int pycd_host_free(void *ptr);

/* NVTX ranges (SURVEY section 5 tracing hook).  Every compute entry point below opens its own range
 * ("pycd.ewald_rows", "pycd.ewald_expand", "pycd.kmc_system_create", "pycd.kmc_advance", "pycd.msd");
 * the drivers bracket their phases (material_setup / material_run / material_msd) with these two. */
int pycd_nvtx_push(const char *name);
int pycd_nvtx_pop(void);

/* ---- Ewald site-pair array --------------------------------------------- */
/* Replaces System.pot_r_ewald / pot_k_ewald / get_precomputed_array
 * (PyCD/core.py:799-878, 1594-1602, 1659-1661).  Pair vectors are never
 * materialised: minimum images (core.py:304-361) are taken on the fly from the
 * site coordinates, the reciprocal sum uses per-site structure factors. */
typedef struct {
    int64_t n_sites;       /* N */
    const double *coords;  /* (N,3) site coordinates, bohr (Material.generate_sites, core.py:154-241) */
    double cell[9];        /* simulation_cell_matrix rows, core.py:296 */
    int32_t pbc[3];        /* core.py:262 */
    double recip[9];       /* System.reciprocal_lattice_matrix rows, core.py:715-722 */
    double volume;         /* System.system_volume, core.py:711-714 */
    double alpha;          /* yaml alpha / ANG2BOHR, core.py:768-771 */
    double r_cut;          /* yaml r_cut * ANG2BOHR, core.py:773-776 */
    double k_cut;          /* yaml k_cut / ANG2BOHR, core.py:779-784 */
    double dielectric;     /* Material.dielectric_constant */
    int32_t k_max[3];      /* ceil(k_cut/|b_i|), core.py:1599 */
    int32_t k_part;        /* with k_parts > 1 (the unit-cell rows sharded over GPUs by k range): this call sums
                              only part k_part of k_parts contiguous parts of the k list; part 0 also carries
                              the real-space and self terms, so the SUM of the parts' outputs (one all-reduce)
                              is P.  k_parts <= 1: everything */
    int64_t plan_rows;     /* 0: the launch plan (tile shape, split-k factor) follows the rows of THIS call.
                              > 0: plan as if plan_rows rows were evaluated.  Ranks that shard an array by row
                              blocks pass the row count of the whole array, so that every block -- and the
                              one-GPU evaluation of all rows -- sums the k vectors in the same order and the
                              all-gathered array is bit-identical to the one-GPU array. */
    int32_t k_parts;
    int32_t reserved;
} pycd_ewald_desc;

typedef struct {
    int64_t k_eff;         /* half-space k vectors with |k|^2 < k_cut^2 */
    int64_t rows;          /* rows computed by this call */
    double fourier_ms;     /* device time of the reciprocal-space kernel */
    double finish_ms;      /* device time of real-space + self + combine */
    double flops;          /* algorithmic flops of this call: 4*rows*N*k_eff + 40*rows*N */
    int32_t k_split;       /* split-k factor used */
} pycd_ewald_stats;

/* out[(row_end-row_begin), N] = P[row_begin:row_end, :].  Row blocks are how the
 * array is sharded across GPUs (rank g owns rows [gN/G,(g+1)N/G)). */
int pycd_ewald_rows(pycd_ctx *ctx, const pycd_ewald_desc *desc, int64_t row_begin,
                    int64_t row_end, double *out, pycd_ewald_stats *stats);

/* The rows of unit cell 0, out[n_basis, N] = P[0:n_basis, :], for a fully periodic supercell whose sites are
 * unit cell 0 plus lattice translations (site = cell * n_basis + basis, cell = (x*ny + y)*nz + z): the
 * reciprocal-space sum of core.py:853-878 factorised by residue classes of the k indices modulo the supercell
 * size -- one pass over the k vectors for the n_basis^2 basis pairs, then a Fourier transform over the
 * n_cells classes (csrc/ewald_cells.cu): n_cells times fewer flops than pycd_ewald_rows(0, n_basis), same sum.
 * Fails with "not translation-invariant ..." when the coordinates / matrices do not have that structure
 * (callers then use pycd_ewald_rows). */
int pycd_ewald_unit_rows(pycd_ctx *ctx, const pycd_ewald_desc *desc, int32_t n_basis, const int32_t size[3],
                         double *out, pycd_ewald_stats *stats);

/* Translation-symmetric expansion (valid for pbc = [1,1,1]): given
 * Pu[n_basis, N] = P[0:n_basis, :] (the rows of unit cell 0) produce
 * out[(row_end-row_begin), N] with P[i,j] = Pu[b_i, (cell_j - cell_i mod size)*n_basis + b_j].
 * Site index layout core.py:402-418 (cell = (x*ny + y)*nz + z, basis fastest). */
int pycd_ewald_expand(pycd_ctx *ctx, const double *p_unit, int32_t n_basis,
                      const int32_t size[3], int64_t row_begin, int64_t row_end, double *out);

/* ---- KMC --------------------------------------------------------------- */
/* Static tables of one carrier type; replaces the per-step lookups of
 * Run.get_process_attributes / get_process_rates (core.py:1946-2050) into
 * Run.__init__'s lists (core.py:1861-1927). */
typedef struct {
    int64_t n_sites;            /* N */
    const double *P;            /* (N,N) precomputed_array; a DEVICE pointer is used in place
                                   (caller keeps it alive), a HOST pointer is copied */
    int64_t n_centres;          /* sites of the carrier's element type */
    const int32_t *site_centre; /* (N) element_type_element_index or -1, core.py:1935-1944 */
    const int32_t *site_class;  /* (N) system_class_index_list, core.py:729-730 */
    int32_t n_class;
    int32_t nn;                 /* System.num_neighbors of the species, core.py:746-765 */
    const int32_t *neigh;       /* (n_centres, nn) new site per slot, core.py:1964-1978 */
    const double *hopvec;       /* (n_centres, nn, 3) hop vectors, bohr, core.py:2037-2040 */
    const double *lam;          /* (n_class, nn) lambda per slot, core.py:1912-1917 */
    const double *vab;          /* (n_class, nn) V_AB per slot, core.py:1918-1922 */
    const double *e_rel;        /* (N) system_relative_energies, core.py:1717-1730 */
    const double *q_lat;        /* (N) base_charge_config, core.py:2530-2538; V_lat = P.q_lat
                                   is computed on the device */
    double q_carrier;           /* species charge, core.py:1830-1832 */
    double kT;                  /* core.py:1705 */
    double vn;                  /* core.py:103 */
    double field[3];            /* core.py:1749-1761 */
    int32_t field_active;
    int32_t p_layout;           /* PYCD_P_DENSE: P is (N,N).  PYCD_P_UNIT_ROWS: P is (n_basis,N) =
                                   the rows of unit cell 0 (pycd_ewald_rows(0, n_basis)); elements are
                                   addressed through lattice translation, pbc = [1,1,1] only */
    int32_t n_basis;            /* UNIT_ROWS: sites per unit cell (<= 255) */
    int32_t size[3];            /* UNIT_ROWS: supercell size (<= 255 per axis) */
} pycd_kmc_system_desc;

#define PYCD_P_DENSE 0
#define PYCD_P_UNIT_ROWS 1

int pycd_kmc_system_create(pycd_ctx *ctx, const pycd_kmc_system_desc *desc, pycd_kmc_system **out);
int pycd_kmc_system_destroy(pycd_kmc_system *sys);
/* copy V_lat back (tests): N values (dense) or n_basis values (unit rows) */
int pycd_kmc_system_vlat(pycd_kmc_system *sys, double *v_lat);

#define PYCD_RNG_REPLAY 0 /* u1,u2 pairs supplied by the host (the reference's random.random(), core.py:2799,2802) */
#define PYCD_RNG_PHILOX 1 /* Philox4x32-10, key = seed, counter = (step, global trajectory id) */

typedef struct {
    int64_t n_traj;            /* trajectories owned by this GPU */
    uint64_t traj_id0;         /* global id of the first one (RNG key; results independent of sharding) */
    int32_t n_carriers;        /* C */
    const int32_t *occupancy0; /* (n_traj, C) initial sites, core.py:2478-2528 */
    double dt_grid;            /* time_interval, a.u., core.py:1710 */
    int64_t n_path;            /* int(t_final/time_interval)+1, core.py:2657 */
    int64_t step_limit;        /* >0: stop a trajectory after this many KMC steps */
    int32_t stop_at_grid_end;  /* 1: reference loop condition end_path_index < n_path (core.py:2787) */
    int32_t rng_mode;
    uint64_t seed;
    int32_t refresh_interval;  /* 1: stateless (every rate re-gathered each step, the reference formulation);
                                  R>1: incremental delta-E updates with a full re-gather every R steps */
    const double *kT_traj;     /* NULL or (n_traj): per-trajectory temperature (sweeps) */
    const double *field_traj;  /* NULL or (n_traj,3): per-trajectory field (implies field_active) */
    int32_t record_unwrapped;  /* 1: keep the (n_traj, n_path, 3C) displacement grid on the device */
    const double *energy0;     /* NULL, or (n_traj) system energy of the initial state (core.py:2778-2780):
                                  turns on the energy / delg_0 grids of output_data */
    /* Doping hooks of Run.do_kmc_steps (core.py:2723-2776); all NULL / 0 for an undoped run.  The
     * dopant distribution itself (site_indices.npy) comes from the reference's material_preprod. */
    const double *e_rel_traj;  /* NULL or (n_traj, N): per-trajectory system_relative_energies with the
                                  shell shifts of relative_energies['doping'] applied (core.py:2750-2764) */
    int32_t n_dopant_max;      /* D: columns of the two arrays below */
    const int32_t *dopant_site;/* NULL or (n_traj, D): dopant sites, -1 = unused entry */
    const double *dopant_dq;   /* (n_traj, D): dopant charge minus the undoped lattice charge of the site
                                  (charge_config, core.py:2553-2557); the device adds
                                  sum_d dq_d P[s, d] to the trajectory's lattice potential */
    const double *dt_grid_traj;/* NULL or (n_traj): per-trajectory time_interval in place of dt_grid (a sweep
                                  flattened into one ensemble: every condition keeps the time grid its own
                                  simulation_parameters.yml would give it) */
} pycd_kmc_ensemble_desc;

int pycd_kmc_ensemble_create(pycd_kmc_system *sys, const pycd_kmc_ensemble_desc *desc,
                             pycd_kmc_ensemble **out);
int pycd_kmc_ensemble_destroy(pycd_kmc_ensemble *ens);
/* Re-arm an ensemble (same sizes and run parameters) with new initial sites: t = 0, empty
 * grid; avoids re-allocating the per-trajectory state between batches.  Stream-ordered: the sites are
 * validated and staged before the call returns (the caller's buffer is free again), the copies and
 * clears are enqueued behind the launches issued so far -- no host synchronisation, so the re-arm and
 * the launches of the next batch can be issued while the current batch still runs. */
int pycd_kmc_ensemble_reset(pycd_kmc_ensemble *ens, const int32_t *occupancy0, uint64_t traj_id0);

/* Advance every unfinished trajectory by at most max_steps KMC steps: rate evaluation
 * (core.py:1989-2050), selection + time advance + recording (core.py:2796-2861).
 * draws: REPLAY mode, (n_traj, 2*max_steps) u1,u2,...; ignored for PHILOX.
 * events_out / times_out: NULL or (n_traj, max_steps): selected process index and
 * simulation time after each step of this call; entries past a trajectory's end hold -1 / 0.0
 * in HOST buffers and are left untouched in DEVICE buffers.  steps_done: NULL or (n_traj) steps
 * taken in this call.
 * n_active: number of trajectories still unfinished after the call. */
int pycd_kmc_advance(pycd_kmc_ensemble *ens, int64_t max_steps, const double *draws,
                     int32_t *events_out, double *times_out, int64_t *steps_done,
                     int64_t *n_active);

/* Batch production without a host round trip per launch (counter-based RNG modes, no per-step outputs):
 * pycd_kmc_advance_async only enqueues the step kernel on the context's stream; any number of calls may be
 * in flight, interleaved with pycd_ctx_flush_l2 / pycd_kmc_read_begin / pycd_kmc_ensemble_reset (all
 * stream-ordered).  pycd_kmc_wait blocks until they have finished, reports the unfinished trajectories and
 * accounts the kernel times (pycd_ctx_total_kernel_ms).  The kernel shape is chosen from the number of
 * unfinished trajectories known at the last synchronising call.  The reference loop has no counterpart:
 * its trajectories run one after the other on the host (core.py:2787-2861). */
int pycd_kmc_advance_async(pycd_kmc_ensemble *ens, int64_t max_steps);
int pycd_kmc_wait(pycd_kmc_ensemble *ens, int64_t *n_active);

/* State read-back; any pointer may be NULL.  unwrapped: (n_traj, n_path, 3C) displacement
 * from the start site on the time grid (unwrapped_traj.npy, core.py:2852-2854);
 * drift: (n_traj, C, 3) sum of hop_vector*k_selected (core.py:2822-2825). */
int pycd_kmc_read(pycd_kmc_ensemble *ens, double *unwrapped, int64_t *n_steps, double *sim_time,
                  int32_t *occupancy, double *drift, int64_t *near_tie, int64_t *clamped,
                  double *rates /* (n_traj, C*nn) rates of the last evaluated step */);
/* Pipelined read-back for batch production (HOST buffers, ideally page-locked; any may be NULL):
 * pycd_kmc_read_begin snapshots the state reached by the launches issued so far and copies it out on
 * a second stream while the caller re-arms the ensemble (pycd_kmc_ensemble_reset, REQUIRED before the
 * next pycd_kmc_advance when `unwrapped` was requested: the ensemble switches to its second displacement
 * grid) and runs the next batch.  pycd_kmc_read_end blocks until the buffers are complete.  One read in
 * flight per ensemble; not available with energy outputs. */
int pycd_kmc_read_begin(pycd_kmc_ensemble *ens, double *unwrapped, int64_t *n_steps, double *sim_time,
                        int32_t *occupancy, double *drift, int64_t *near_tie, int64_t *clamped);
int pycd_kmc_read_end(pycd_kmc_ensemble *ens);
/* output_data 'energy' and 'delg_0' (core.py:2807-2809, 2855-2857): (n_traj, n_path) each; either may be NULL */
int pycd_kmc_read_energy(pycd_kmc_ensemble *ens, double *energy_grid, double *dg0_grid);
/* device pointer of the resident displacement grid (input of pycd_msd without a copy) */
int pycd_kmc_unwrapped_device(pycd_kmc_ensemble *ens, const double **dev_ptr);
/* Diagnostics: name of the step kernel the last pycd_kmc_advance launched (NUL-terminated,
 * truncated to n bytes), and whether the UNIT_ROWS system carries the lattice-stencil tables
 * (one table entry H[b_a][cell_y - cell_a][b_y][slot] = P[n_slot(a), y] - P[a, y] per carrier
 * pair: the row difference of core.py:2004-2008 precomputed under translation symmetry).
 * why: reason when unavailable. */
int pycd_kmc_last_kernel(pycd_kmc_ensemble *ens, char *buf, int32_t n);
int pycd_kmc_system_stencil(pycd_kmc_system *sys, int32_t *available, int64_t *table_bytes, char *why,
                            int32_t n);

/* ---- MSD --------------------------------------------------------------- */
/* Replaces the lag loop of Analysis.compute_msd (core.py:2996-3022):
 * sd[traj,tau,c] = mean_t |r(t+tau,c)-r(t,c)|^2 * scale^2, averaged over the carriers
 * of each type.  unwrapped: (n_traj, n_path, 3C) bohr; scale = dist_conversion;
 * type_offsets: (n_types+1) carrier index ranges; sd_species: (n_traj, n_msd, n_types);
 * sd_carrier: NULL or (n_traj, n_msd, C). */
int pycd_msd(pycd_ctx *ctx, const double *unwrapped, int64_t n_traj, int64_t n_path,
             int32_t n_carriers, int64_t n_msd, double scale, const int32_t *type_offsets,
             int32_t n_types, double *sd_species, double *sd_carrier);

#ifdef __cplusplus
}
#endif
#endif /* PYCD_B200_H */
