"""Worker of tests/test_gpu_multi.py (one process per rank, launched by torch.distributed.run): the
multi-GPU data path on real devices -- row-sharded Ewald array + all-gather, trajectory-sharded KMC
ensemble, MSD partial sums reduced at the end -- each compared with the one-GPU result computed by the
same process.  PYCD_DIST_BACKEND=nccl (one GPU per rank, NVLink) or gloo (ranks share GPU 0; the
collectives then run on host tensors)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / 'oracle', ROOT / 'tests'):
    sys.path.insert(0, str(p))


def main():
    rank, world, local_rank = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    backend = os.environ.get('PYCD_DIST_BACKEND', 'nccl')
    device = local_rank if backend == 'nccl' else 0
    torch.cuda.set_device(device)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if backend == 'nccl':
        dist.init_process_group('nccl', device_id=torch.device('cuda', device))
    else:
        dist.init_process_group('gloo')
    import helpers as H
    from pycd_b200 import _native as nat
    from pycd_b200 import constants
    from pycd_b200 import dist as D
    from pycd_b200 import ewald as EW
    from pycd_b200 import kmc as K
    from pycd_b200 import msd as M
    from pycd_b200.kmc import RunParameters
    from pycd_b200.lattice import Supercell
    ctx = nat.default_context(device)
    dev = torch.device('cuda', device)
    comm_dev = dev if backend == 'nccl' else torch.device('cpu')

    ex = H.load_example('hematite')
    sim = ex.sim
    sc = Supercell(ex.lattice, [3, 3, 2], [1, 1, 1])
    n = sc.num_system_elements
    run = RunParameters(ex.lattice, sc, sc.hop_neighbor_tables(), sim['temp'], 'full', 'full', sim['t_final'],
                        sim['time_interval'], [24, 0], {}, sim['relative_energies'], sim['external_field'])
    ep = EW.EwaldParameters(sc, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])
    coords = torch.from_numpy(np.ascontiguousarray(sc.coordinates)).to(dev)

    # 1. Ewald: rank g evaluates rows [gN/G, (g+1)N/G) directly (dense formulation, BASELINE config 4),
    #    one all-gather completes the array on every rank; it must equal the one-GPU array
    for n_rows in (n, n - 7):   # equal and ragged row blocks
        full = torch.zeros((n_rows, n), dtype=torch.float64, device=dev)
        lo, hi = D.block(rank, world, n_rows)
        EW.ewald_rows(ctx, ep, coords.data_ptr(), lo, hi, out=full[lo:hi].data_ptr(), plan_rows=n_rows)
        torch.cuda.synchronize()
        if backend == 'nccl':
            D.allgather_rows(full)
        else:
            host = full.cpu()
            D.allgather_rows(host)
            full = host.to(dev)
        single = torch.empty((n_rows, n), dtype=torch.float64, device=dev)
        EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, n_rows, out=single.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(full, single), f'rank {rank}: gathered Ewald array differs from the one-GPU array'

    # 1b. rows of unit cell 0 sharded by k range: every rank sums one part of the k list, one all-reduce
    part = torch.empty((sc.n_per_cell, n), dtype=torch.float64, device=dev)
    EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, sc.n_per_cell, out=part.data_ptr(), k_part=rank, k_parts=world)
    torch.cuda.synchronize()
    if backend == 'nccl':
        dist.all_reduce(part)
    else:
        host = part.cpu()
        dist.all_reduce(host)
        part = host.to(dev)
    one = torch.empty((sc.n_per_cell, n), dtype=torch.float64, device=dev)
    EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, sc.n_per_cell, out=one.data_ptr())
    torch.cuda.synchronize()
    assert float((part - one).abs().max()) <= 1e-13 * float(one.abs().max()), 'k-range shards do not add up'

    # 2. KMC: trajectories in contiguous blocks, Philox keyed by the global id; sharded == unsharded
    p_unit = torch.empty((sc.n_per_cell, n), dtype=torch.float64, device=dev)
    EW.ewald_rows(ctx, ep, coords.data_ptr(), 0, sc.n_per_cell, out=p_unit.data_ptr())
    system = K.KmcSystem(ctx, run, p_unit.data_ptr(), layout='unit_rows')
    n_traj, steps, n_path = 10, 2048, 64
    kw = dict(dt_grid=run.time_interval / 3000, n_path=n_path, step_limit=steps, stop_at_grid_end=False,
              rng_mode=nat.RNG_PHILOX, seed=17, refresh_interval=64)
    lo, hi = D.block(rank, world, n_traj)
    occ_all = K.philox_initial_occupancy(run.tables, n_traj, 24, 17)

    def run_block(a, b):
        ens = K.KmcEnsemble(system, occ_all[a:b], traj_id0=a, **kw)
        while ens.advance_resident(1024) > 0:
            pass
        out = ens.read()
        kernel = ens.last_kernel()
        ens.close()
        return out, kernel
    mine, kernel = run_block(lo, hi)
    whole, _ = run_block(0, n_traj)
    assert kernel.startswith('kmc_step_warp_kernel'), kernel
    assert np.array_equal(mine['unwrapped'], whole['unwrapped'][lo:hi])
    assert np.array_equal(mine['occupancy'], whole['occupancy'][lo:hi])
    gathered = D.gather_trajectory_arrays(mine['unwrapped'], n_traj, device=comm_dev)
    assert np.array_equal(gathered, whole['unwrapped'])

    # 3. MSD: per-rank species-averaged SD, {sum, sum of squares} all-reduced == one-GPU analysis
    toff = np.array([0, 24], dtype=np.int32)
    n_msd = 33
    avg_mine = M.species_avg_sd(ctx, mine['unwrapped'], hi - lo, n_path, 24, n_msd, 1 / constants.ANG2BOHR, toff)
    avg_all = M.species_avg_sd(ctx, whole['unwrapped'], n_traj, n_path, 24, n_msd, 1 / constants.ANG2BOHR, toff)
    part = torch.tensor(np.stack([avg_mine[:, :, 0].sum(0), (avg_mine[:, :, 0] ** 2).sum(0)]), device=comm_dev)
    dist.all_reduce(part)
    mean = (part[0] / n_traj).cpu().numpy()
    assert np.allclose(mean, avg_all[:, :, 0].mean(0), rtol=1e-13, atol=0)
    full_avg = D.gather_trajectory_arrays(avg_mine, n_traj, device=comm_dev)
    assert np.array_equal(full_avg, avg_all)
    system.close()
    dist.barrier()
    if rank == 0:
        Path(sys.argv[1]).write_text(f'ok world={world} backend={backend} kernel={kernel}')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
