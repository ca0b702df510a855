"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-block all-gather of the
Ewald array, trajectory sharding with globally keyed initial states, and the MSD reduction
that must equal the single-process analysis."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, tmp):
    for p in (ROOT, ROOT / 'oracle', ROOT / 'tests'):
        sys.path.insert(0, str(p))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import helpers as H
    import oracle as O
    from pycd_b200 import dist as D
    from pycd_b200 import kmc as K
    from pycd_b200 import msd as M
    from pycd_b200 import constants

    # 1. Ewald row blocks (computed by the CPU oracle here) all-gathered into the full array
    ex = H.load_example('hematite')
    n = ex.supercell.num_system_elements
    for n_rows in (n, n - 1):  # equal and ragged blocks
        full = torch.zeros((n_rows, n), dtype=torch.float64)
        lo, hi = D.block(rank, world, n_rows)
        full[lo:hi] = torch.from_numpy(ex.P[lo:hi])
        D.allgather_rows(full)
        assert np.array_equal(full.numpy(), ex.P[:n_rows])

    # 2. trajectories sharded by contiguous blocks, keyed by global id
    ex4 = H.load_example('hematite', species_count=[4, 0])
    run = H.run_parameters(ex4)
    n_traj = 6
    lo, hi = D.block(rank, world, n_traj)
    occ_all = K.philox_initial_occupancy(run.tables, n_traj, 4, seed=9)
    occ_mine = K.philox_initial_occupancy(run.tables, hi - lo, 4, seed=9, traj_id0=lo)
    assert np.array_equal(occ_all[lo:hi], occ_mine)
    kw = dict(dt_grid=run.time_interval / 20, n_path=64, step_limit=1200, stop_at_grid_end=False,
              rng_mode=1, seed=9)
    orc = O.KmcOracle(run, ex4.P, **kw)
    mine = orc.ensemble(occ_mine, traj_id0=lo)
    whole = orc.ensemble(occ_all, traj_id0=0)
    assert np.array_equal(mine['unwrapped'], whole['unwrapped'][lo:hi])

    # 3. MSD reduction: gather per-rank species-averaged SD, analyse, compare with 1 process
    mp_ = M.MsdParameters(3, [4, 0], n_traj, 64 * 1e-9, 1e-9, 20.0, 3, 300)
    assert mp_.n_path == 65 or mp_.n_path == 64
    n_path = 64
    mp_.n_path = n_path

    def species_avg(uw):
        pos = uw.reshape(uw.shape[0], n_path, 4, 3) * mp_.dist_conversion
        return O.msd_sd(pos, mp_.n_msd).mean(axis=2, keepdims=True)
    avg_all = D.gather_trajectory_arrays(species_avg(mine['unwrapped']), n_traj)
    res = M.analyse(mp_, avg_all)
    ref = M.analyse(mp_, species_avg(whole['unwrapped']))
    assert np.array_equal(res['msd_data'], ref['msd_data'])
    assert np.array_equal(res['diffusivity'], ref['diffusivity'])
    assert np.array_equal(res['diffusivity_sem'], ref['diffusivity_sem'])
    if rank == 0:
        Path(tmp, 'ok').write_text('ok')
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok').exists()


def test_block_partition():
    from pycd_b200.dist import block
    for n, w in ((30000, 8), (17, 4), (3, 8)):
        blocks = [block(r, w, n) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
