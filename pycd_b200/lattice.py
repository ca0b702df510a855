"""Material parameters, supercell sites, minimum images and hop-neighbour tables.

Host-side geometry that feeds the CUDA path.  Semantics follow the reference's
`Material` (PyCD/core.py:37-241), `Neighbors` (:244-668) and the cell-geometry part of
`System.__init__` (:709-726), but everything is vectorised numpy and the neighbour
table exploits lattice translation, so the 10x10x10 supercells of the benchmark
configs (which the reference cannot set up: O(N^2) Python loops) take
seconds.
"""
import itertools
import warnings
from collections import defaultdict
from types import SimpleNamespace

import numpy as np

from . import constants
from .fileio import read_poscar


class Lattice:
    """Unit cell + material parameters (reference: Material, core.py:42-152).

    Atoms are ordered per element by ascending fractional z with a STABLE sort; the
    reference's default argsort breaks ties platform-dependently.
    """

    def __init__(self, params):
        pos = read_poscar(params.input_coord_file_location)
        self.element_types = list(pos['element_types'])
        self.n_elements_per_unit_cell = np.asarray(pos['num_elements'], dtype=int)
        self.total_elements_per_unit_cell = int(pos['total_elements'])
        lattice_ang = np.array(pos['lattice_matrix'], dtype=float)
        frac = pos['coordinates']
        if pos['coordinate_type'].lower().startswith('c'):  # Cartesian, core.py:62-64
            frac = np.dot(frac, np.linalg.inv(lattice_ang))
        self.lattice_matrix = lattice_ang * constants.ANG2BOHR  # core.py:65
        self.num_element_types = len(self.element_types)
        self.element_type_index_list = np.repeat(np.arange(self.num_element_types),
                                                 self.n_elements_per_unit_cell)
        self.name = params.name
        self.species_types = list(params.species_types)
        self.num_species_types = len(self.species_types)
        self.species_charge_list = params.species_charge_list
        self.species_to_element_type_map = dict(params.species_to_element_type_map)
        self.class_list_input = params.class_list

        self.fractional_unit_cell_coords = np.zeros_like(frac)
        self.unit_cell_class_list = []
        head = 0
        for ti, name in enumerate(self.element_types):
            block = frac[self.element_type_index_list == ti]
            order = np.argsort(block[:, 2], kind='stable')  # core.py:90-96 (stable: F11)
            n = len(block)
            self.fractional_unit_cell_coords[head:head + n] = block[order]
            self.unit_cell_class_list.extend(int(params.class_list[name][i]) - 1 for i in order)
            head += n
        self.cartesian_unit_cell_coords = np.dot(self.fractional_unit_cell_coords,
                                                 self.lattice_matrix)  # core.py:99-100
        self.charge_types = params.charge_types
        self.vn = params.vn / constants.SEC2AUTIME  # core.py:103
        ev = constants.EV2HARTREE
        self.lambda_values = {k: [[v * ev for v in row] for row in rows]
                              for k, rows in params.lambda_values.items()}
        self.v_ab = {k: [[v * ev for v in row] for row in rows] for k, rows in params.v_ab.items()}
        a2b = constants.ANG2BOHR
        self.neighbor_cutoff_dist = {k: [[(v * a2b) if v else None for v in row] for row in rows]
                                     for k, rows in params.neighbor_cutoff_dist.items()}
        self.neighbor_cutoff_dist_tol = {k: [[(v * a2b) if v else None for v in row] for row in rows]
                                         for k, rows in params.neighbor_cutoff_dist_tol.items()}
        self.element_type_delimiter = params.element_type_delimiter
        self.dielectric_constant = params.dielectric_constant
        self.num_classes = [len(set(params.class_list[name])) for name in self.element_types]
        self.element_type_to_species_map = defaultdict(list)
        for name in self.element_types:
            for sp in self.species_types:
                if name in self.species_to_element_type_map[sp]:
                    self.element_type_to_species_map[name].append(sp)
        d = self.element_type_delimiter
        self.hop_element_types = {
            sp: [d.join(c) for c in itertools.product(self.species_to_element_type_map[sp], repeat=2)]
            for sp in self.species_types}

    def element_head(self, element_type_index):
        """Offset of the element's first atom inside a unit cell."""
        return int(self.n_elements_per_unit_cell[:element_type_index].sum())


class Supercell:
    """size[0] x size[1] x size[2] replication of a Lattice (reference: Neighbors +
    the geometry attributes of System).  Site index = cell*n_per_cell + local with
    cell = (x*ny + y)*nz + z (core.py:402-418, pinned by the reference's test_neighbors.py)."""

    def __init__(self, lattice, system_size, pbc):
        self.lattice = lattice
        self.system_size = np.asarray(system_size, dtype=int)
        self.pbc = np.asarray(pbc, dtype=int)
        self.n_dim = 3
        self.num_cells = int(self.system_size.prod())
        self.n_per_cell = lattice.total_elements_per_unit_cell
        self.num_system_elements = self.num_cells * self.n_per_cell
        sx, sy, sz = (int(v) for v in self.system_size)
        gx, gy, gz = np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing='ij')
        self.cell_indices = np.column_stack((gx.ravel(), gy.ravel(), gz.ravel()))
        shifts = np.dot(self.cell_indices, lattice.lattice_matrix)  # core.py:214
        self.coordinates = (lattice.cartesian_unit_cell_coords[None, :, :]
                            + shifts[:, None, :]).reshape(-1, 3)  # core.py:218-222
        # minimum-image cell: rows scaled (core.py:296)
        self.cell_matrix = lattice.lattice_matrix * self.system_size[:, None]
        self.cell_matrix_inv = np.linalg.inv(self.cell_matrix)
        # Ewald cell: np.multiply(size, L) scales COLUMNS (core.py:709-710)
        self.translational_matrix = np.multiply(self.system_size, lattice.lattice_matrix)
        t = self.translational_matrix
        self.system_volume = abs(np.dot(t[0], np.cross(t[1], t[2])))
        self.reciprocal_lattice_matrix = (2 * np.pi / self.system_volume
                                          * np.array([np.cross(t[1], t[2]), np.cross(t[2], t[0]),
                                                      np.cross(t[0], t[1])]))
        self.translational_vector_length = np.linalg.norm(t, axis=1)
        self.reciprocal_lattice_vector_length = np.linalg.norm(self.reciprocal_lattice_matrix, axis=1)
        if not np.allclose(self.translational_matrix, self.cell_matrix, rtol=1e-12, atol=1e-12):
            warnings.warn('supercell: the reference scales lattice COLUMNS for the Ewald cell and '
                          'ROWS for minimum images; they differ for this size/lattice (SURVEY F10), '
                          'replicated literally')
        self.system_class_index_list = np.tile(np.asarray(lattice.unit_cell_class_list, dtype=np.int32),
                                               self.num_cells)  # core.py:729-730

    # -- indices -------------------------------------------------------------------------
    def element_sites(self, element_type_index):
        """All sites of an element in ascending site index (core.py:2508-2523)."""
        lat = self.lattice
        head = lat.element_head(element_type_index)
        n_t = int(lat.n_elements_per_unit_cell[element_type_index])
        return (np.arange(self.num_cells)[:, None] * self.n_per_cell + head
                + np.arange(n_t)[None, :]).ravel()

    def site_centre_table(self, element_type_index):
        """site -> element_type_element_index (core.py:1935-1944) or -1."""
        out = np.full(self.num_system_elements, -1, dtype=np.int32)
        sites = self.element_sites(element_type_index)
        out[sites] = np.arange(len(sites), dtype=np.int32)
        return out

    # -- minimum image -------------------------------------------------------------------
    def _image_offsets(self):
        """Candidate lattice offsets in the reference's order: baseline first, then
        dx, dy, dz ascending (core.py:333-348)."""
        offs = [(0, 0, 0)]
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    if dx == 0 and dy == 0 and dz == 0:
                        continue
                    d = (dx if self.pbc[0] else 0, dy if self.pbc[1] else 0, dz if self.pbc[2] else 0)
                    if d == (0, 0, 0):
                        continue
                    offs.append(d)
        return np.array(offs, dtype=float)

    def min_image(self, disp):
        """Vectorised apply_minimum_image_convention (core.py:304-361) for disp[..., 3]."""
        disp = np.asarray(disp, dtype=float)
        shape = disp.shape
        d = disp.reshape(-1, 3)
        f = d @ self.cell_matrix_inv
        base = np.zeros_like(f)
        m = self.pbc.astype(bool)
        base[:, m] = -np.round(f[:, m])
        best = None
        best_n = None
        for o in self._image_offsets():  # strict '<' keeps the first minimum, like the reference
            cand = (f + (base + o)) @ self.cell_matrix
            n = np.sqrt((cand * cand).sum(axis=1))
            if best is None:
                best, best_n = cand, n
            else:
                upd = n < best_n
                best = np.where(upd[:, None], cand, best)
                best_n = np.where(upd, n, best_n)
        return best.reshape(shape)

    def pairwise_min_image_vectors(self, max_sites=4096):
        """(N,N,3) array of the reference's pairwise_min_image_vector_data.npy
        (core.py:571-594); only for small cells (24*N^2 bytes)."""
        n = self.num_system_elements
        if n > max_sites:
            raise ValueError(f'pairwise vector file for N={n} would take {24 * n * n / 1e9:.1f} GB; '
                             'the CUDA path never needs it')
        out = np.empty((n, n, 3))
        for i in range(n):
            out[i] = self.min_image(self.coordinates - self.coordinates[i])
        return out

    # -- hop neighbour tables ------------------------------------------------------------
    def hop_neighbor_tables(self, force_all_pairs=False):
        """Dense neighbour tables of every hop key:
            tables[key][class][hop_dist] = SimpleNamespace(index [n_centres, nn_h] int64 site
            indices ascending per row, vector [n_centres, nn_h, 3] bohr)
        Selection rule of the reference (core.py:529-533): lo < |min image| <= hi over all
        sites of the element, listed in ascending site index."""
        lat = self.lattice
        tables = {}
        for key, class_rows in lat.neighbor_cutoff_dist.items():
            centre_name = key.split(lat.element_type_delimiter)[0]
            ti = lat.element_types.index(centre_name)
            n_t = int(lat.n_elements_per_unit_cell[ti])
            head = lat.element_head(ti)
            sites = self.element_sites(ti)
            n_centres = len(sites)
            full_pbc = bool(np.all(self.pbc == 1))
            translate = full_pbc and not force_all_pairs and n_centres > 256
            rows_sites = sites[:n_t] if translate else sites  # unit cell 0 only when translating
            disp = self.min_image(self.coordinates[sites][None, :, :]
                                  - self.coordinates[rows_sites][:, None, :])
            dist = np.sqrt((disp * disp).sum(axis=2))
            tol = lat.neighbor_cutoff_dist_tol[key]
            max_hop = max(v for row in class_rows for v in row if v)
            if np.any(self.pbc):
                half = 0.5 * min(np.linalg.norm(self.cell_matrix[k]) for k in range(3) if self.pbc[k])
                if max_hop >= half - 1e-9:
                    warnings.warn(f'{key}: hop distance {max_hop / constants.ANG2BOHR:.4f} A reaches '
                                  'half the supercell; +L/2 and -L/2 images coincide and hop-vector '
                                  'signs are floating-point coin flips (SURVEY F12)')
            key_tables = []
            for ci, row in enumerate(class_rows):
                cls_tables = []
                for hi, cutoff in enumerate(row):
                    lo, hi_lim = cutoff - tol[ci][hi], cutoff + tol[ci][hi]
                    mask = (dist > lo) & (dist <= hi_lim)
                    counts = mask.sum(axis=1)
                    width = int(counts.max()) if counts.size else 0
                    # neighbour columns in ascending order, padded with -1 (counts differ between
                    # site classes; only the rows of the list's own class are used at run time)
                    order = np.argsort(~mask, axis=1, kind='stable')[:, :width]
                    valid = np.take_along_axis(mask, order, axis=1)
                    cols = np.where(valid, order, -1)
                    vec = np.where(valid[:, :, None],
                                   disp[np.arange(len(rows_sites))[:, None], order], 0.0)
                    if translate:
                        idx, vec, counts = self._translate_rows(cols, vec, counts, n_t, head)
                    else:
                        idx = np.where(valid, sites[order], -1)
                    cls_tables.append(SimpleNamespace(index=idx.astype(np.int64),
                                                      vector=np.ascontiguousarray(vec),
                                                      count=np.asarray(counts, dtype=np.int64)))
                key_tables.append(cls_tables)
            tables[key] = key_tables
        return tables

    def _translate_rows(self, cols, vec, counts, n_t, head):
        """Neighbour lists of unit cell 0 (cols = centre indices cell*n_t + local, -1 = pad)
        -> lists of every cell, re-sorted by ascending site index like the reference."""
        sx, sy, sz = (int(v) for v in self.system_size)
        width = cols.shape[1]
        pad = cols < 0
        safe = np.where(pad, 0, cols)
        ncell = safe // n_t
        nloc = safe % n_t
        nz = ncell % sz
        ny = (ncell // sz) % sy
        nx = ncell // (sz * sy)
        cx, cy, cz = self.cell_indices[:, 0], self.cell_indices[:, 1], self.cell_indices[:, 2]
        tx = (cx[:, None, None] + nx[None]) % sx
        ty = (cy[:, None, None] + ny[None]) % sy
        tz = (cz[:, None, None] + nz[None]) % sz
        site = ((tx * sy + ty) * sz + tz) * self.n_per_cell + head + nloc[None]
        big = np.iinfo(np.int64).max
        site = np.where(pad[None], big, site).reshape(self.num_cells * n_t, width)
        v = np.broadcast_to(vec[None], (self.num_cells,) + vec.shape).reshape(
            self.num_cells * n_t, width, 3)
        order = np.argsort(site, axis=1, kind='stable')
        site = np.take_along_axis(site, order, axis=1)
        v = np.take_along_axis(v, order[:, :, None], axis=1)
        site = np.where(site == big, -1, site)
        return site, v, np.tile(counts, self.num_cells)
