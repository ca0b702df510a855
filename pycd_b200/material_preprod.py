"""Drop-in for PyCD/material_preprod.py:13 restricted to what the undoped path needs:
the per-trajectory MT19937 state files of Run.preproduction (core.py:2573-2580)."""
from .config import load_simulation_parameters
from .kmc import write_initial_rnd_states


def material_preprod(dst_path):
    sim = load_simulation_parameters(dst_path)
    if sim.get('doping') and any(sim['doping'].get('num_dopants', [])):
        raise NotImplementedError('doping is outside the accelerated path (SURVEY section 2)')
    write_initial_rnd_states(dst_path, int(sim['n_traj']), sim['random_seed'])
    return None
