"""Doping hooks of the KMC run: what `Run.do_kmc_steps` does with a trajectory's
`site_indices.npy` before the first step (PyCD/core.py:2723-2776).

The dopant DISTRIBUTION (random / pairwise / gradient insertion, shell bookkeeping,
core.py:2084-2476, written by the reference's material_preprod) is control plane and out of
scope; the run consumes its product.  `site_indices.npy` rows are
`[site index, dopant element type index, dopant element index, shell index]`
(core.py:2424-2476): shell 0 rows are the dopant sites themselves, higher shells their
neighbours.  Per trajectory the hooks

  * shift `system_relative_energies[site]` by `relative_energies['doping'][substituted
    element][dopant type][shell]` (eV) for every listed site (core.py:2750-2764);
  * replace the lattice charge of each dopant site by `doping['charge'][ion_charge_type]
    [dopant element]` (`charge_config`, core.py:2553-2557);
  * optionally start the carriers ON dopant sites (`site_charge_initiation`,
    `generate_initial_occupancy`, core.py:2486-2496).

On the device the charge replacement is a per-trajectory correction of the lattice
potential, `V_lat'[s] = V_lat[s] + sum_d (q_d - q_lat[d]) P[s, d]`, and the energy shift a
per-trajectory copy of the site-energy table (csrc/kmc.cu `vlat_doped_kernel`).
"""
import numpy as np

from . import constants


class TrajectoryDoping:
    """Doping state of ONE trajectory: dopant sites per dopant element, the shifted site
    energies and the charge replacements."""

    def __init__(self, dopant_site_indices, e_rel, sites, dq):
        self.dopant_site_indices = dopant_site_indices   # {dopant element: [site, ...]} (shell 0 rows, file order)
        self.e_rel = e_rel                                # (N) system_relative_energies of the trajectory
        self.sites = np.asarray(sites, dtype=np.int32)   # dopant sites, one entry per replaced charge
        self.dq = np.asarray(dq, dtype=np.float64)       # dopant charge minus the undoped lattice charge

    def q_lat(self, base):
        """charge_config before the carriers are added (core.py:2551-2557)."""
        q = np.array(base, dtype=np.float64, copy=True)
        q[self.sites] += self.dq
        return q


class DopingHooks:
    """Static part (Run.__init__, core.py:1732-1736, 1763-1781)."""

    def __init__(self, lattice, doping, relative_energies, ion_charge_type):
        self.cfg = doping or {}
        self.active = bool(np.any(self.cfg.get('num_dopants', [0])))
        self.lattice = lattice
        self.ion_charge_type = ion_charge_type
        self.dopant_element_types, self.substitution_element_types, self.dopant_species_types = [], [], []
        if not self.active:
            return
        for entry in self.cfg['doping_element_map']:
            sub, dop = entry.split(lattice.element_type_delimiter)
            self.dopant_element_types.append(dop)
            self.substitution_element_types.append(sub)
            self.dopant_species_types.append(lattice.element_type_to_species_map[sub][0])
        self.shell_energies = (relative_energies or {}).get('doping', {})

    @property
    def pairwise_insertion(self):
        """core.py:2726-2735: one site_indices.npy for the whole run instead of one per trajectory."""
        types = self.cfg.get('insertion_type', [])
        if 'pairwise' in types:
            return self.cfg['num_dopants'][types.index('pairwise')] != 0
        return False

    def site_indices_path(self, traj_dir):
        return (traj_dir.parent if self.pairwise_insertion else traj_dir) / 'site_indices.npy'

    def load(self, site_indices, undoped_e_rel, undoped_q_lat):
        """site_indices: (n, 4) int array -> TrajectoryDoping."""
        si = np.asarray(site_indices)
        if si.ndim != 2 or si.shape[1] != 4:
            raise ValueError(f'site_indices must have shape (n, 4), not {si.shape}')
        dopant_sites = {}
        for row in si[si[:, 3] == 0]:   # core.py:2741-2748
            dopant_sites.setdefault(self.dopant_element_types[int(row[1])], []).append(int(row[0]))
        e_rel = np.array(undoped_e_rel, dtype=np.float64, copy=True)
        for site, type_index, _, shell in si:   # core.py:2750-2764 (one add per row, file order)
            sub = self.substitution_element_types[int(type_index)]
            table = self.shell_energies[sub][int(type_index)]
            if shell < len(table):
                e_rel[int(site)] += table[int(shell)] * constants.EV2HARTREE
        sites, dq = [], []
        for dop, where in dopant_sites.items():   # charge_config, core.py:2553-2557 (assignment, not +=)
            q_d = float(self.cfg['charge'][self.ion_charge_type][dop])
            for s in dict.fromkeys(where):
                sites.append(s)
                dq.append(q_d - float(undoped_q_lat[s]))
        return TrajectoryDoping(dopant_sites, e_rel, sites, dq)

    def initiation_sites(self, rng, species_type, n_species, dopant_site_indices):
        """Carriers started on dopant sites (core.py:2486-2496).  Returns (sites, carriers left).
        Like the reference, `rng.sample` raises if there are fewer dopant sites than carriers."""
        occ = []
        for map_index, dopant_species_type in enumerate(self.dopant_species_types):
            n_dopant_sites = self.cfg['num_dopants'][map_index]
            if (self.cfg['site_charge_initiation'][map_index] == 'yes' and dopant_species_type == species_type
                    and n_dopant_sites and n_species):
                where = dopant_site_indices[self.dopant_element_types[map_index]]
                occ.extend(rng.sample(where, n_species))
                n_species -= len(where[:n_species])
        return occ, n_species
