"""The two YAML files of a PyCD work directory (reference: PyCD/material_run.py:15-53,
material_setup.py:21-34).  Parsed with yaml.safe_load (the reference's bare yaml.load
needs PyYAML < 6).  New, optional knobs of this implementation live under the `b200:`
key of simulation_parameters.yml so that existing files load unchanged:

    b200:                      # simulation_parameters.yml (material_run)
      rng: replay | philox     # default replay = the reference's MT19937 streams
      refresh_interval: 256    # philox mode; 1 = stateless rate evaluation (the reference formulation,
                               # always used in replay mode)
      chunk_steps: 32768       # KMC steps per kernel launch
      p_layout: auto           # auto | unit_rows | dense: which Ewald table the step kernel reads

    b200:                      # sys_config.yml (material_setup)
      ewald_symmetric: auto    # auto | true | false: rows of unit cell 0 + translation, or all rows directly
      dense_limit_gb: 4        # the dense N x N file is skipped above this size when the unit rows exist
"""
from types import SimpleNamespace

import numpy as np
import yaml


def load_yaml(path):
    with open(path, 'r') as fh:
        return yaml.safe_load(fh)


def load_simulation_parameters(dst_path):
    sim = load_yaml(dst_path / 'simulation_parameters.yml')
    sim['system_size'] = np.asarray(sim['system_size'])
    sim['pbc'] = np.asarray(sim['pbc'])
    sim['species_count'] = np.asarray(sim['species_count'])
    sim.setdefault('b200', {})
    return sim


def input_directory(dst_path, sim, parallel_aware=True):
    """InputFiles location from work_dir_depth (+1 in parallel mode), material_run.py:31-40."""
    depth = sim['work_dir_depth']
    if parallel_aware and sim.get('compute_mode') == 'parallel':
        depth += 1
    base = dst_path.resolve()
    if depth != 0:
        base = base.parents[depth - 1]
    return base / sim['input_file_directory_name']


def load_material_parameters(input_directory_path):
    params = load_yaml(input_directory_path / 'sys_config.yml')
    params['input_coord_file_location'] = input_directory_path / 'POSCAR'
    return SimpleNamespace(**params)
