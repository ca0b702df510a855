"""Drop-in for PyCD/material_setup.py:13-85: neighbour list, pair vectors and the Ewald
precomputed array, with the N^2*K arithmetic on the GPU."""
from datetime import datetime

import numpy as np

from . import _native as nat
from . import ewald as ew
from .config import load_material_parameters
from .fileio import generate_report
from .lattice import Lattice, Supercell
from .tables import save_hop_neighbor_list


def material_setup(input_directory_path, system_size, pbc, generate_hop_neighbor_list,
                   generate_pairwise_min_image_vector_data, generate_precomputed_array,
                   compute_energy_contributions, return_k_vector_data):
    """Same signature and files as the reference.  `return_k_vector_data` (a per-k energy
    ranking report, core.py:1632-1653) is a diagnostic outside the hot path and raises."""
    params = load_material_parameters(input_directory_path)
    lattice = Lattice(params)
    supercell = Supercell(lattice, np.asarray(system_size), np.asarray(pbc))

    if generate_hop_neighbor_list:
        start = datetime.now()
        input_directory_path.mkdir(parents=True, exist_ok=True)
        save_hop_neighbor_list(input_directory_path / 'hop_neighbor_list.npy',
                               supercell.hop_neighbor_tables())
        generate_report(start, input_directory_path, 'neighbor_list', 1)

    if generate_pairwise_min_image_vector_data:
        np.save(input_directory_path / 'pairwise_min_image_vector_data.npy',
                supercell.pairwise_min_image_vectors())

    if generate_precomputed_array:
        if return_k_vector_data:
            raise NotImplementedError('return_k_vector_data: the per-k energy report is not part '
                                      'of the accelerated path')
        start = datetime.now()
        ep = ew.EwaldParameters(supercell, params.alpha, params.r_cut, params.k_cut)
        ctx = nat.default_context()
        P, _ = ew.precomputed_array(ctx, ep)
        energies = None
        if compute_energy_contributions:
            energies = energy_contributions(ctx, ep, P)
        generate_report(start, input_directory_path, 'precomputed_array', 1,
                        ''.join(ew.log_prefix(ep, energies)))
        np.save(input_directory_path / 'precomputed_array.npy', P)
    return None


def energy_contributions(ctx, ep, P):
    """sum_ij q_i q_j X_ij for X = real, fourier, self with 'full' ion charges
    (core.py:1663-1674).  The real-space part comes from a second pass with an empty
    k list (k_cut -> 0), the reciprocal part by difference."""
    q = ep.base_charges('full')
    total = float(q @ P @ q)
    self_term = -np.sqrt(ep.alpha / np.pi) / ep.dielectric
    self_e = float(np.dot(q, q) * self_term)
    k_cut, k_max = ep.k_cut, ep.k_max.copy()
    try:
        ep.k_cut, ep.k_max = 0.0, np.zeros(3, dtype=int)
        real_plus_self, _ = ew.precomputed_array(ctx, ep)
    finally:
        ep.k_cut, ep.k_max = k_cut, k_max
    real_e = float(q @ real_plus_self @ q) - self_e
    return real_e, total - real_e - self_e, self_e
