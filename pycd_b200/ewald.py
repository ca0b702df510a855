"""Host side of the Ewald precompute: parameters, the C-ABI call, the log file.

Mirrors `System.__init__` (Ewald part, PyCD/core.py:767-784) and
`System.get_precomputed_array` (core.py:1611-1679); the N^2*K arithmetic itself
(core.py:799-878) runs in csrc/ewald.cu.
"""
import ctypes as C

import numpy as np

from . import _native as nat
from . import constants


class EwaldParameters:
    """alpha / r_cut / k_cut in atomic units with the reference's conversions
    (core.py:768-784: alpha and k_cut are divided by ANG2BOHR, r_cut multiplied).
    Only user-specified floats are supported; the 'optimal' / 'converge' /
    'simulation_cell' searches (core.py:913-1592) are outside the hot path."""

    def __init__(self, supercell, alpha, r_cut, k_cut):
        for name, v in (('alpha', alpha), ('r_cut', r_cut), ('k_cut', k_cut)):
            if isinstance(v, (str, list)) or not np.isreal(v):
                raise NotImplementedError(
                    f"{name}={v!r}: only user-specified float Ewald parameters are supported "
                    "(the reference's parameter searches are out of scope, SURVEY section 2)")
        self.supercell = supercell
        self.alpha = alpha / constants.ANG2BOHR
        self.r_cut = r_cut * constants.ANG2BOHR
        self.k_cut = k_cut / constants.ANG2BOHR
        self.dielectric = supercell.lattice.dielectric_constant
        # core.py:1599-1600
        self.k_max = np.ceil(self.k_cut / supercell.reciprocal_lattice_vector_length).astype(int)
        self.num_k_vectors = np.ceil(np.prod(2 * self.k_max + 1) * np.pi / 6 - 1).astype(int)

    def desc(self, coords):
        sc = self.supercell
        d = nat.EwaldDesc()
        d.n_sites = sc.num_system_elements
        d.coords = nat.ptr(coords)
        d.cell[:] = sc.cell_matrix.ravel().tolist()
        d.pbc[:] = [int(v) for v in sc.pbc]
        d.recip[:] = sc.reciprocal_lattice_matrix.ravel().tolist()
        d.volume = float(sc.system_volume)
        d.alpha = float(self.alpha)
        d.r_cut = float(self.r_cut)
        d.k_cut = float(self.k_cut)
        d.dielectric = float(self.dielectric)
        d.k_max[:] = [int(v) for v in self.k_max]
        return d

    def base_charges(self, ion_charge_type='full'):
        """base_charge_config_for_accuracy_analysis, core.py:941-948."""
        lat = self.supercell.lattice
        unit = np.array([lat.charge_types[ion_charge_type][lat.element_types[i]]
                         for i in lat.element_type_index_list], dtype=float)
        return np.tile(unit, self.supercell.num_cells)

    def cutoff_error_lines(self):
        """The two 'cutoff error' log lines, core.py:977-984."""
        sc = self.supercell
        q = self.base_charges('full')
        q2 = float(np.dot(q, q))
        alpha, r_cut, k_cut = self.alpha, self.r_cut, self.k_cut
        length = np.power(sc.system_volume, 1 / 3)
        real = q2 * np.sqrt(r_cut / (2 * sc.system_volume)) * (np.exp(-(alpha * r_cut) ** 2)
                                                               / (alpha * r_cut) ** 2)
        n_cut = k_cut * length / (2 * np.pi)
        x = np.pi * n_cut / (alpha * length)
        fourier = q2 * np.sqrt(n_cut) / (alpha * length ** 2) * (np.exp(-x ** 2) / x ** 2)
        return [f'Real-space cutoff error: {real:.3e}\n',
                f'Fourier-space cutoff error: {fourier:.3e}\n\n']


def ewald_rows(ctx, params, coords, row_begin, row_end, out=None, plan_rows=0, k_part=0, k_parts=1):
    """P[row_begin:row_end, :] through pycd_ewald_rows.  coords / out: numpy arrays (host
    path) or torch CUDA tensors / raw device addresses (resident path).  plan_rows: row count of
    the whole array when this call evaluates one block of a row-sharded array (every block then
    uses the launch plan of the one-GPU evaluation and the gathered array is bit-identical to it).
    k_part / k_parts: sum only one contiguous part of the k list (part 0 carries the real-space and self
    terms too): the unit-cell rows sharded over GPUs by k range, completed by ONE all-reduce (sum)."""
    n = params.supercell.num_system_elements
    if out is None:
        out = np.empty((row_end - row_begin, n))
    desc = params.desc(coords)
    desc.plan_rows = int(plan_rows)
    desc.k_part, desc.k_parts = int(k_part), int(k_parts)
    stats = nat.EwaldStats()
    nat.check(nat.lib().pycd_ewald_rows(ctx.handle, C.byref(desc), int(row_begin), int(row_end),
                                        nat.ptr(out), C.byref(stats)))
    return out, {'k_eff': stats.k_eff, 'rows': stats.rows, 'fourier_ms': stats.fourier_ms,
                 'finish_ms': stats.finish_ms, 'flops': stats.flops, 'k_split': stats.k_split}


def unit_cell_rows(ctx, params, coords=None, out=None, method='auto'):
    """P[0:n_per_cell, :] of a fully periodic supercell -- what the step kernel reads (unit-rows layout) and what
    the translation expansion starts from.  method 'cells' = the class-factorised sum (pycd_ewald_unit_rows:
    one pass over the k vectors for the n_per_cell^2 basis pairs + a Fourier transform over the cells),
    'rows' = pycd_ewald_rows(0, n_per_cell) (the DMMA kernel: 4 n_per_cell N K_eff flop), 'auto' = cells when
    the geometry has the structure it needs, else rows."""
    sc = params.supercell
    n = sc.num_system_elements
    if coords is None:
        coords = np.ascontiguousarray(sc.coordinates)
    if method not in ('auto', 'cells', 'rows'):
        raise ValueError(f"method must be auto, cells or rows, not {method!r}")
    if method != 'rows':
        if out is None:
            out = np.empty((sc.n_per_cell, n))
        desc = params.desc(coords)
        stats = nat.EwaldStats()
        size = (C.c_int32 * 3)(*[int(v) for v in sc.system_size])
        rc = nat.lib().pycd_ewald_unit_rows(ctx.handle, C.byref(desc), int(sc.n_per_cell), size, nat.ptr(out),
                                            C.byref(stats))
        if rc == 0:
            return out, {'k_eff': stats.k_eff, 'rows': stats.rows, 'fourier_ms': stats.fourier_ms,
                         'finish_ms': stats.finish_ms, 'flops': stats.flops, 'k_split': stats.k_split,
                         'method': 'cells'}
        msg = nat.lib().pycd_last_error().decode(errors='replace')
        if method == 'cells' or 'not translation-invariant' not in msg and 'pbc' not in msg:
            raise nat.NativeError(msg)
    out, st = ewald_rows(ctx, params, coords, 0, sc.n_per_cell, out=out)
    st['method'] = 'rows'
    return out, st


def ewald_expand(ctx, supercell, p_unit, row_begin, row_end, out=None):
    """Rows of the full array from the unit-cell-0 rows (pycd_ewald_expand)."""
    n = supercell.num_system_elements
    if out is None:
        out = np.empty((row_end - row_begin, n))
    size = (C.c_int32 * 3)(*[int(v) for v in supercell.system_size])
    nat.check(nat.lib().pycd_ewald_expand(ctx.handle, nat.ptr(p_unit), int(supercell.n_per_cell), size,
                                          int(row_begin), int(row_end), nat.ptr(out)))
    return out


def can_use_translation_symmetry(supercell):
    return bool(np.all(supercell.pbc == 1)) and supercell.num_cells > 1


def precomputed_array(ctx, params, coords=None, symmetric=None, row_begin=0, row_end=None, out=None):
    """The (rows of the) N x N precomputed array.

    symmetric=True computes only the rows of unit cell 0 (n_per_cell x N) and expands by
    lattice translation (exact for full PBC: P depends on (basis_i, basis_j, cell_j - cell_i));
    symmetric=False evaluates every requested row directly.  Default: symmetric when valid."""
    sc = params.supercell
    n = sc.num_system_elements
    if coords is None:
        coords = np.ascontiguousarray(sc.coordinates)
    if row_end is None:
        row_end = n
    if symmetric is None:
        symmetric = can_use_translation_symmetry(sc)
    if symmetric and not can_use_translation_symmetry(sc):
        raise ValueError('translation symmetry needs pbc = [1, 1, 1] and more than one cell')
    if not symmetric:
        out, stats = ewald_rows(ctx, params, coords, row_begin, row_end, out)
        stats['symmetric'] = False
        return out, stats
    p_unit, stats = unit_cell_rows(ctx, params, coords)
    out = ewald_expand(ctx, sc, p_unit, row_begin, row_end, out)
    stats['symmetric'] = True
    stats['expand_ms'] = ctx.last_kernel_ms(nat.KC_EWALD_EXPAND)
    return out, stats


def log_prefix(params, energies=None):
    """Lines of precomputed_array.log (core.py:1578-1585, 1656-1657, 1663-1674).  Line 3
    ('alpha: ...') is what material_run parses (material_run.py:75-80).  The reference's
    first two lines are wall-clock ratios of its own benchmark_ewald (core.py:913-939);
    there is nothing to time here, so they carry a fixed 0 and stay line-compatible."""
    a2b = constants.ANG2BOHR
    lines = [f'tau_ratio, (tau_r/tau_f): {0.0:.3e}\n',
             f'time_ratio, (time_r/time_f): {0.0:.3e}\n\n',
             f'alpha: {params.alpha * a2b:.3e} / angstrom (user-specified)\n',
             f'r_cut: {params.r_cut / a2b:.3e} angstrom (user-specified)\n',
             f'k_cut: {params.k_cut * a2b:.3e} / angstrom (user-specified)\n']
    lines += params.cutoff_error_lines()
    k = params.k_max
    lines.append(f'k_max: [{k[0]}, {k[1]}, {k[2]}]\n')
    lines.append(f'number of k-vectors: {params.num_k_vectors}\n\n')
    if energies is not None:
        ev = constants.EV2HARTREE
        real, fourier, self_e = energies
        lines.append(f'Energy contribution from Real space: {real / ev} eV\n')
        lines.append(f'Energy contribution from Fourier-space: {fourier / ev} eV\n')
        lines.append(f'Energy contribution from self-interactions: {self_e / ev} eV\n')
        lines.append(f'Total system energy (neutral): {(real + fourier + self_e) / ev} eV\n\n')
    return lines
