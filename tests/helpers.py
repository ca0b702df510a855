"""Shared loaders for the golden fixtures (tests/golden/, built by make_golden.py)."""
import io
import pickle
import random
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import yaml

from pycd_b200.ewald import EwaldParameters
from pycd_b200.kmc import RunParameters
from pycd_b200.lattice import Lattice, Supercell
from pycd_b200.tables import load_hop_neighbor_list

GOLD = Path(__file__).resolve().parent / 'golden'


def load_example(name, species_count=None, sim_override=None):
    """name in {'hematite', 'bvo'} -> namespace(lattice, supercell, sim, sys_cfg, hop, P, ...)."""
    d = GOLD / name
    sim = yaml.safe_load(open(d / 'simulation_parameters.yml'))
    if sim_override:
        sim = sim_override
    if species_count is not None:
        sim['species_count'] = list(species_count)
    cfg = yaml.safe_load(open(d / 'InputFiles' / 'sys_config.yml'))
    cfg['input_coord_file_location'] = d / 'InputFiles' / 'POSCAR'
    lat = Lattice(SimpleNamespace(**cfg))
    sc = Supercell(lat, sim['system_size'], sim['pbc'])
    hop = load_hop_neighbor_list(d / 'InputFiles' / 'hop_neighbor_list.npy')
    P = np.load(d / 'InputFiles' / 'precomputed_array.npy')
    return SimpleNamespace(dir=d, sim=sim, cfg=cfg, lattice=lat, supercell=sc, hop=hop, P=P)


def run_parameters(ex, sim=None):
    sim = sim or ex.sim
    return RunParameters(ex.lattice, ex.supercell, ex.hop, sim['temp'], sim['ion_charge_type'],
                         sim['species_charge_type'], sim['t_final'], sim['time_interval'],
                         sim['species_count'], sim['initial_occupancy'], sim['relative_energies'],
                         sim['external_field'], sim.get('doping'))


def ewald_parameters(ex):
    return EwaldParameters(ex.supercell, ex.cfg['alpha'], ex.cfg['r_cut'], ex.cfg['k_cut'])


def shipped_pairwise(ex):
    return np.load(ex.dir / 'InputFiles' / 'pairwise_min_image_vector_data.npy')


def shipped_coords(ex):
    """Site coordinates in the SHIPPED site order: X_j := pairwise[0, j]."""
    return np.ascontiguousarray(shipped_pairwise(ex)[0])


def rng_from_state_bytes(raw):
    r = random.Random()
    r.setstate(pickle.load(io.BytesIO(bytes(raw))))
    return r


def shipped_rng(ex):
    return rng_from_state_bytes((ex.dir / 'traj1' / 'initial_rnd_state.dump').read_bytes())


def shipped_unwrapped(ex):
    return np.load(ex.dir / 'traj1' / 'unwrapped_traj.npz')['unwrapped']


def shipped_time_sample(ex):
    z = np.load(ex.dir / 'traj1' / 'time_data_sample.npz')
    return int(z['n']), z['index'], z['value']


def draw_stream(rng, n_steps):
    rnd = rng.random
    return np.array([rnd() for _ in range(2 * n_steps)])


def load_ref_case(tag):
    """ref_<tag>.npz written by make_golden.py from a run of the unmodified reference."""
    z = np.load(GOLD / f'ref_{tag}.npz', allow_pickle=False)
    sim = yaml.safe_load(str(z['sim_yaml']))
    ex = load_example(str(z['example']).lower(), sim_override=sim)
    return ex, z
