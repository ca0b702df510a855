"""Hop-neighbour list I/O (the reference's pickled `hop_neighbor_list.npy`) and the flat
per-carrier-type tables the step kernel reads.

Reference layout (PyCD/core.py:565-569, 660-663): a pickled dict
    {hop key: [class][hop_dist] -> ReturnValues(neighbor_system_element_indices = object
     array of int arrays, displacement_vector_list = object array of (nn,3) f64 [bohr],
     num_neighbors = int array)}
whose class is `PyCD.core.ReturnValues`; reading must not need the reference package and
writing must stay loadable by it.
"""
import contextlib
import pickle
import sys
import types

import numpy as np


class ReturnValues(object):
    """Attribute bag; stands in for PyCD.core.ReturnValues when (un)pickling."""

    def __init__(self, **kwargs):
        for key, value in kwargs.items():
            setattr(self, key, value)


ReturnValues.__module__ = 'PyCD.core'
ReturnValues.__qualname__ = 'ReturnValues'


class _HopListUnpickler(pickle.Unpickler):
    """hop_neighbor_list.npy is a pickled dict of ReturnValues objects holding numpy arrays
    (core.py:565-569, 663): only those globals are resolved; anything else in the file is refused
    instead of being imported and called."""
    _NUMPY = {('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
              ('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'),
              ('numpy', 'ndarray'), ('numpy', 'dtype'),
              ('numpy.core.numeric', '_frombuffer'), ('numpy._core.numeric', '_frombuffer'),
              ('_codecs', 'encode'), ('__builtin__', 'bytes'), ('builtins', 'bytes')}   # byte strings of protocol-2 pickles

    def find_class(self, module, name):
        if name == 'ReturnValues':
            return ReturnValues
        if (module, name) in self._NUMPY:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f'hop_neighbor_list.npy must not reference {module}.{name}')


def load_hop_neighbor_list(path):
    """Reads hop_neighbor_list.npy without importing the reference."""
    with open(path, 'rb') as fh:
        version = np.lib.format.read_magic(fh)
        if version == (1, 0):
            np.lib.format.read_array_header_1_0(fh)
        else:
            np.lib.format.read_array_header_2_0(fh)
        obj = _HopListUnpickler(fh).load()
    if isinstance(obj, np.ndarray):
        obj = obj[()]
    return obj


@contextlib.contextmanager
def _reference_class_visible():
    """Makes `PyCD.core.ReturnValues` resolvable while pickling so that the file can be
    loaded by the reference (pickle stores the class by module path)."""
    created = []
    if 'PyCD' not in sys.modules:
        sys.modules['PyCD'] = types.ModuleType('PyCD')
        created.append('PyCD')
    if 'PyCD.core' not in sys.modules:
        mod = types.ModuleType('PyCD.core')
        sys.modules['PyCD.core'] = mod
        created.append('PyCD.core')
    core = sys.modules['PyCD.core']
    had = getattr(core, 'ReturnValues', None)
    core.ReturnValues = ReturnValues
    try:
        yield
    finally:
        if had is not None:
            core.ReturnValues = had
        for name in created:
            sys.modules.pop(name, None)


def tables_to_hop_neighbor_list(tables):
    """Dense tables of Supercell.hop_neighbor_tables -> the reference's nested structure."""
    out = {}
    for key, class_tables in tables.items():
        key_list = []
        for cls_tables in class_tables:
            cls_list = []
            for t in cls_tables:
                n = len(t.count)
                idx = np.empty(n, dtype=object)
                vec = np.empty(n, dtype=object)
                for i in range(n):
                    c = int(t.count[i])
                    idx[i] = np.asarray(t.index[i, :c], dtype=np.int64)
                    vec[i] = np.asarray(t.vector[i, :c], dtype=float)
                cls_list.append(ReturnValues(neighbor_system_element_indices=idx,
                                             displacement_vector_list=vec,
                                             num_neighbors=np.asarray(t.count, dtype=int)))
            key_list.append(cls_list)
        out[key] = key_list
    return out


def save_hop_neighbor_list(path, tables):
    """Writes hop_neighbor_list.npy in the reference's pickle layout (core.py:663)."""
    hop = tables_to_hop_neighbor_list(tables) if _is_dense(tables) else tables
    with _reference_class_visible():
        with open(path, 'wb') as fh:
            np.save(fh, hop, allow_pickle=True)


def _is_dense(tables):
    first = next(iter(tables.values()))
    return hasattr(first[0][0], 'count')


class HopTables:
    """Flat tables of ONE carrier type: what Run.__init__ (core.py:1861-1927)
    and get_process_attributes (core.py:1946-1987) look up per step.

    neigh[e, slot], hopvec[e, slot, :] for e = element_type_element_index; slots are
    hop-distance types in yaml order, neighbours ascending by site index inside each; the
    lists used for a site are those of the site's own class (core.py:1961-1970).
    """

    def __init__(self, lattice, supercell, hop_neighbor_list, species_type):
        lat = lattice
        self.species_type = species_type
        self.species_index = lat.species_types.index(species_type)
        self.hop_key = lat.hop_element_types[species_type][0]  # core.py:1871
        element = lat.species_to_element_type_map[species_type][0]
        self.element_type_index = lat.element_types.index(element)
        ti = self.element_type_index
        n_t = int(lat.n_elements_per_unit_cell[ti])
        head = lat.element_head(ti)
        self.sites = supercell.element_sites(ti)
        self.n_centres = len(self.sites)
        self.site_centre = supercell.site_centre_table(ti)
        self.site_class = np.ascontiguousarray(supercell.system_class_index_list, dtype=np.int32)
        self.n_class = int(lat.num_classes[ti])
        hop = hop_neighbor_list[self.hop_key]
        dense = hasattr(hop[0][0], 'count')
        n_hop = len(lat.neighbor_cutoff_dist[self.hop_key][0])

        def rows(ci, hi):
            t = hop[ci][hi]
            if dense:
                return t.index, t.vector, t.count
            return t.neighbor_system_element_indices, t.displacement_vector_list, t.num_neighbors

        # num_neighbors of the species: class 0, site 0 (core.py:746-765)
        self.nn = int(sum(int(rows(0, hi)[2][0]) for hi in range(n_hop)))
        centre_class = self.site_class[self.sites]
        # the reference picks the first unit-cell site of each class as sample (core.py:1842-1859)
        # and assumes every site of the class has the same per-distance counts
        self.neigh = np.full((self.n_centres, self.nn), -1, dtype=np.int32)
        self.hopvec = np.zeros((self.n_centres, self.nn, 3))
        self.lam = np.zeros((self.n_class, self.nn))
        self.vab = np.zeros((self.n_class, self.nn))
        self.slot_hop_dist = np.zeros((self.n_class, self.nn), dtype=np.int32)
        for ci in range(self.n_class):
            members = np.nonzero(centre_class == ci)[0]
            if len(members) == 0:
                continue
            slot = 0
            for hi in range(len(lat.neighbor_cutoff_dist[self.hop_key][ci])):
                idx, vec, cnt = rows(ci, hi)
                c = int(cnt[members[0]])
                if not np.all(np.asarray(cnt)[members] == c):
                    raise ValueError(f'{self.hop_key}: class {ci} sites disagree on the number of '
                                     f'neighbours at hop distance {hi}')
                if slot + c > self.nn:
                    raise ValueError(f'{self.hop_key}: class {ci} has more neighbour slots than class 0 '
                                     '(the reference assumes equal counts, core.py:750-765)')
                if dense:
                    self.neigh[members, slot:slot + c] = idx[members, :c]
                    self.hopvec[members, slot:slot + c] = vec[members, :c]
                else:
                    for m in members:
                        self.neigh[m, slot:slot + c] = np.asarray(idx[m])
                        self.hopvec[m, slot:slot + c] = np.asarray(vec[m]).reshape(c, 3)
                self.lam[ci, slot:slot + c] = lat.lambda_values[self.hop_key][ci][hi]
                self.vab[ci, slot:slot + c] = lat.v_ab[self.hop_key][ci][hi]
                self.slot_hop_dist[ci, slot:slot + c] = hi
                slot += c
            if slot != self.nn:
                raise ValueError(f'{self.hop_key}: class {ci} fills {slot} of {self.nn} neighbour slots')
        if (self.neigh < 0).any():
            raise ValueError('incomplete neighbour table')
        self.n_type_per_cell = n_t
        self.head = head
