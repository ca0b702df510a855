"""The multi-GPU path on hardware: gathered Ewald array == one-GPU array, sharded ensembles
== unsharded, reduced MSD == one-GPU MSD.  `nccl`: one GPU per rank (skipped below two GPUs); `gloo`: the
same worker with both ranks on GPU 0, so that a one-GPU box still runs the sharded code on a device."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize('backend,n_ranks', [('gloo', 2), ('nccl', 2), ('nccl', 4)])
def test_sharded_paths_equal_single_gpu(tmp_path, backend, n_ranks):
    import torch
    if backend == 'nccl' and torch.cuda.device_count() < n_ranks:
        pytest.skip(f'needs {n_ranks} GPUs')
    ok = tmp_path / 'ok'
    env = dict(os.environ, PYCD_DIST_BACKEND=backend)
    port = 29900 + os.getpid() % 90
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n_ranks}',
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          str(ROOT / 'tests' / 'multi_gpu_worker.py'), str(ok)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert ok.read_text().startswith(f'ok world={n_ranks} backend={backend}')
