"""Drop-in for PyCD/material_setup.py:13-85: neighbour list, pair vectors and the Ewald
precomputed array, with the N^2*K arithmetic on the GPU."""
import os
from datetime import datetime

import numpy as np

from . import _native as nat
from . import ewald as ew
from .config import load_material_parameters
from .fileio import generate_report
from .lattice import Lattice, Supercell
from .tables import save_hop_neighbor_list

# rows of unit cell 0, (n_per_cell, N) f64: written for pbc = [1, 1, 1], read by material_run
UNIT_ROWS_FILE = 'precomputed_array_unit_rows.npy'


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
        ctx = nat.default_context(int(os.environ.get('LOCAL_RANK', '0')))
        opts = getattr(params, 'b200', None) or {}
        n = supercell.num_system_elements
        symmetric = ew.can_use_translation_symmetry(supercell)
        # b200.ewald_symmetric (sys_config.yml): auto = use translation symmetry when valid; false = evaluate
        # all N rows directly (the reference's formulation; cfg 4); true = insist on it
        want_sym = opts.get('ewald_symmetric', 'auto')
        if want_sym not in ('auto', True, False, 'true', 'false'):
            raise ValueError(f"b200.ewald_symmetric must be auto, true or false, not {want_sym!r}")
        if want_sym in (True, 'true') and not symmetric:
            raise ValueError('b200.ewald_symmetric: true needs pbc = [1, 1, 1] and more than one cell')
        if want_sym in (False, 'false'):
            symmetric = False
        # Full PBC: the rows of unit cell 0 determine the whole array.  They are written
        # next to the reference's file; material_run prefers them (L2-resident table, stencil kernel).
        # The dense N x N file stays the exchange format with the reference and is written unless it would
        # exceed b200.dense_limit_gb (default 4; 7.2 GB at Hematite 10x10x10).
        limit = float(opts.get('dense_limit_gb', 4.0)) * 1e9
        want_dense = (not symmetric) or n * n * 8 <= limit
        P = None
        if symmetric:
            p_unit, _ = ew.unit_cell_rows(ctx, ep, np.ascontiguousarray(supercell.coordinates))
            np.save(input_directory_path / UNIT_ROWS_FILE, p_unit)
            if want_dense:
                P = ew.ewald_expand(ctx, supercell, p_unit, 0, n)
        else:
            P, _ = ew.precomputed_array(ctx, ep, symmetric=False)
        energies = None
        if compute_energy_contributions:
            if P is None:
                raise NotImplementedError('compute_energy_contributions needs the dense array; raise '
                                          'b200.dense_limit_gb in sys_config.yml or pass 0')
            energies = energy_contributions(ctx, ep, P)
        generate_report(start, input_directory_path, 'precomputed_array', 1,
                        ''.join(ew.log_prefix(ep, energies)))
        if P is not None:
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
