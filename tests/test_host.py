"""Host logic that needs no GPU: constants, geometry, tables, file formats, the C-ABI
library's exported symbols (no compute calls)."""
import ctypes
import pickle
import random
import re
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import yaml

import helpers as H
from pycd_b200 import _native as nat
from pycd_b200 import constants
from pycd_b200 import kmc as K
from pycd_b200.ewald import EwaldParameters, log_prefix
from pycd_b200.fileio import format_elapsed, generate_report, read_poscar
from pycd_b200.lattice import Lattice, Supercell
from pycd_b200.tables import HopTables, load_hop_neighbor_list, save_hop_neighbor_list

ROOT = Path(__file__).resolve().parents[1]


def test_constants_pinned():
    # SURVEY appendix A.1 values of the reference's PyCD/constants.py
    assert constants.ANG2BOHR == 1.8897261264678997
    assert constants.EV2HARTREE == 0.03674932246011127
    assert constants.SEC2AUTIME == 4.134137336634339e16
    assert 300 * constants.K2AUTEMP == 9.500431539230842e-4


def test_poscar_and_lattice():
    ex = H.load_example('hematite')
    lat = ex.lattice
    assert lat.element_types == ['Fe', 'O']
    assert list(lat.n_elements_per_unit_cell) == [12, 18]
    z = lat.fractional_unit_cell_coords
    assert np.all(np.diff(z[:12, 2]) >= 0) and np.all(np.diff(z[12:, 2]) >= 0)
    assert lat.vn == 1.85e13 / constants.SEC2AUTIME
    assert lat.lambda_values['Fe:Fe'][0][0] == 1.74533 * constants.EV2HARTREE
    assert lat.hop_element_types == {'electron': ['Fe:Fe'], 'hole': ['O:O']}
    bvo = H.load_example('bvo').lattice
    assert bvo.num_classes == [2, 1, 1]
    # stable z-sort: BVO O classes alternate by z-layer (shipped order)
    assert sorted(set(bvo.unit_cell_class_list[:16])) == [0, 1]


def test_site_index_layout():
    """cell = (x*ny + y)*nz + z, basis fastest (reference tests/test_neighbors.py:54-80)."""
    ex = H.load_example('hematite')
    sc = Supercell(ex.lattice, [3, 2, 4], [1, 1, 1])
    npc = sc.n_per_cell
    for (x, y, z, b) in [(0, 0, 0, 0), (1, 0, 0, 5), (2, 1, 3, 29), (0, 1, 2, 12)]:
        idx = ((x * 2 + y) * 4 + z) * npc + b
        want = ex.lattice.cartesian_unit_cell_coords[b] + np.dot([x, y, z], ex.lattice.lattice_matrix)
        assert np.allclose(sc.coordinates[idx], want, rtol=0, atol=1e-12)
    fe = sc.element_sites(0)
    assert len(fe) == 12 * 24 and fe[12] == npc
    t = sc.site_centre_table(0)
    assert t[npc + 3] == 12 + 3 and t[12] == -1


def test_cell_geometry_matches_reference_formulas():
    ex = H.load_example('hematite')
    sc = ex.supercell
    # size[0]==size[1] and block-diagonal lattice: row and column scaling agree
    assert np.allclose(sc.translational_matrix, sc.cell_matrix)
    ep = H.ewald_parameters(ex)
    assert list(ep.k_max) == [10, 10, 16]
    assert int(ep.num_k_vectors) == 7619
    bvo = H.load_example('bvo')
    assert list(H.ewald_parameters(bvo).k_max) == [9, 9, 10]


def test_min_image_partial_pbc():
    ex = H.load_example('hematite')
    sc = Supercell(ex.lattice, [2, 2, 1], [1, 0, 1])
    d = np.array([[30.0, 25.0, 40.0]])
    out = sc.min_image(d)
    f_in, f_out = d @ sc.cell_matrix_inv, out @ sc.cell_matrix_inv
    assert abs(f_out[0, 1] - f_in[0, 1]) < 1e-12         # non-periodic axis untouched
    assert abs(f_out[0, 0]) <= 0.5 + 1e-12 and abs(f_out[0, 2]) <= 0.5 + 1e-12


@pytest.mark.parametrize('name', ['hematite', 'bvo'])
def test_neighbour_tables_match_shipped_lists(name):
    import warnings
    ex = H.load_example(name)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t = ex.supercell.hop_neighbor_tables()
    for key, classes in ex.hop.items():
        for ci, hops in enumerate(classes):
            for hi, h in enumerate(hops):
                mine = t[key][ci][hi]
                assert np.array_equal(h.num_neighbors, mine.count)
                for i, idx in enumerate(h.neighbor_system_element_indices):
                    assert np.array_equal(np.asarray(idx), mine.index[i, :mine.count[i]])
    if name == 'hematite':  # non-degenerate: vectors too (BVO 2x2x1 is degenerate)
        h = ex.hop['Fe:Fe'][0][0]
        vec = np.array([np.asarray(v) for v in h.displacement_vector_list])
        assert np.abs(vec - t['Fe:Fe'][0][0].vector).max() < 1e-13


def test_translation_path_equals_all_pairs():
    ex = H.load_example('hematite')
    sc = Supercell(ex.lattice, [5, 5, 2], [1, 1, 1])
    a, b = sc.hop_neighbor_tables(), sc.hop_neighbor_tables(force_all_pairs=True)
    for hi in range(2):
        assert np.array_equal(a['Fe:Fe'][0][hi].index, b['Fe:Fe'][0][hi].index)
        assert np.abs(a['Fe:Fe'][0][hi].vector - b['Fe:Fe'][0][hi].vector).max() < 1e-12


def test_full_size_tables_match_brute_force_on_sampled_sites():
    """BASELINE config 3 (Hematite 10x10x10, N = 30 000): the tables the oracle AND the CUDA path read at
    full size come from the translation path; here rows of sampled Fe sites are rebuilt independently --
    literal 27-image minimum image of the site against every Fe site (core.py:304-361) and the
    reference's selection rule lo < r <= hi in ascending site index (core.py:529-533, 617-640)."""
    ex = H.load_example('hematite')
    sc = Supercell(ex.lattice, [10, 10, 10], [1, 1, 1])
    assert sc.num_system_elements == 30000
    tabs = sc.hop_neighbor_tables()['Fe:Fe'][0]
    lat = ex.lattice
    ti = lat.element_types.index('Fe')
    sites = sc.element_sites(ti)
    rng = np.random.default_rng(5)
    picks = np.concatenate([[0, 11, len(sites) - 1], rng.choice(len(sites), size=21, replace=False)])
    cut, tol = lat.neighbor_cutoff_dist['Fe:Fe'][0], lat.neighbor_cutoff_dist_tol['Fe:Fe'][0]
    cell = sc.cell_matrix
    shifts = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], dtype=float) @ cell
    for e in picks:
        d0 = sc.coordinates[sites] - sc.coordinates[sites[e]]                  # (n_centres, 3)
        cand = d0[:, None, :] + shifts[None, :, :]                             # all 27 images, literally
        r = np.sqrt((cand * cand).sum(axis=2))
        best = r.argmin(axis=1)
        rmin = r[np.arange(len(sites)), best]
        vmin = cand[np.arange(len(sites)), best]
        for hi in range(2):
            sel = np.nonzero((rmin > cut[hi] - tol[hi]) & (rmin <= cut[hi] + tol[hi]))[0]
            row = tabs[hi].index[e]
            assert np.array_equal(row[row >= 0], sites[sel]), (e, hi)
            assert np.abs(tabs[hi].vector[e][:len(sel)] - vmin[sel]).max() < 1e-9
    tab = HopTables(lat, sc, sc.hop_neighbor_tables(), 'electron')
    assert tab.nn == 4 and tab.n_centres == 12000


def test_hop_list_roundtrip_in_reference_pickle_format(tmp_path):
    ex = H.load_example('hematite')
    t = ex.supercell.hop_neighbor_tables()
    f = tmp_path / 'hop_neighbor_list.npy'
    save_hop_neighbor_list(f, t)
    raw = f.read_bytes()
    assert b'PyCD.core' in raw and b'ReturnValues' in raw  # loadable by the reference
    back = load_hop_neighbor_list(f)
    h = back['Fe:Fe'][0][1]
    assert np.array_equal(np.asarray(h.neighbor_system_element_indices[5]), t['Fe:Fe'][0][1].index[5, :1])
    assert 'PyCD.core' not in sys.modules or hasattr(sys.modules['PyCD.core'], 'Material')
    tab = HopTables(ex.lattice, ex.supercell, back, 'electron')
    tab2 = HopTables(ex.lattice, ex.supercell, t, 'electron')
    assert np.array_equal(tab.neigh, tab2.neigh) and np.array_equal(tab.hopvec, tab2.hopvec)


def test_flat_tables():
    ex = H.load_example('hematite')
    tab = HopTables(ex.lattice, ex.supercell, ex.hop, 'electron')
    assert tab.nn == 4 and tab.n_centres == 48 and tab.neigh.shape == (48, 4)
    ev = constants.EV2HARTREE
    assert np.allclose(tab.lam[0], np.array([1.74533] * 3 + [1.88683]) * ev)
    assert np.allclose(tab.vab[0], np.array([0.184] * 3 + [0.028]) * ev)
    assert np.all(np.diff(tab.neigh[:, :3], axis=1) > 0)  # ascending inside a hop distance
    bvo = H.load_example('bvo')
    tv = HopTables(bvo.lattice, bvo.supercell, bvo.hop, 'electron')
    assert tv.nn == 6  # degenerate 2x2x1 cell
    th = HopTables(bvo.lattice, bvo.supercell, bvo.hop, 'hole')
    assert th.nn == 12 and th.n_class == 2 and th.n_centres == 64


def test_run_parameters_and_initial_state():
    ex, z = H.load_ref_case('bvo_2h')
    run = H.run_parameters(ex)
    assert run.species_type == 'hole' and run.q_carrier == 1.0 and run.n_proc == 24
    assert set(np.round(run.e_rel / constants.EV2HARTREE, 4)) == {0.0, 0.0406}
    assert run.n_path == int(run.t_final / run.time_interval) + 1
    rng = H.rng_from_state_bytes(z['rnd_state_0'])
    assert run.initial_occupancy_from(rng) == list(z['occ0'])
    with pytest.raises(NotImplementedError):
        H.run_parameters(H.load_example('bvo', species_count=[1, 1]))


def test_preproduction_states_match_shipped_dump(tmp_path):
    ex = H.load_example('hematite')
    K.write_initial_rnd_states(tmp_path, 3, ex.sim['random_seed'])
    mine = pickle.load(open(tmp_path / 'traj1' / 'initial_rnd_state.dump', 'rb'))
    gold = pickle.load(open(ex.dir / 'traj1' / 'initial_rnd_state.dump', 'rb'))
    assert mine == gold
    assert (tmp_path / 'traj3' / 'initial_rnd_state.dump').exists()


def test_precomputed_array_log_is_parseable_like_the_reference(tmp_path):
    ex = H.load_example('hematite')
    ep = H.ewald_parameters(ex)
    from datetime import datetime
    generate_report(datetime.now(), tmp_path, 'precomputed_array', 1, ''.join(log_prefix(ep)))
    lines = open(tmp_path / 'precomputed_array.log').read().splitlines()
    assert float(lines[3][7:16]) == 0.2667          # material_run.py:75-80
    assert lines[3] == 'alpha: 2.667e-01 / angstrom (user-specified)'
    assert lines[4] == 'r_cut: 4.618e+00 angstrom (user-specified)'
    assert lines[5] == 'k_cut: 6.998e+00 / angstrom (user-specified)'
    assert lines[6] == 'Real-space cutoff error: 2.406e+00'
    assert lines[7] == 'Fourier-space cutoff error: 4.450e-76'
    assert lines[9] == 'k_max: [10, 10, 16]' and lines[10] == 'number of k-vectors: 7619'
    assert lines[-1].startswith('Time elapsed:  0 hours,  0 minutes')


def test_ewald_parameter_searches_are_rejected():
    ex = H.load_example('hematite')
    with pytest.raises(NotImplementedError):
        EwaldParameters(ex.supercell, 'optimal', 4.6, 7.0)


def test_philox_initial_occupancy_is_sharding_invariant():
    ex = H.load_example('hematite', species_count=[6, 0])
    run = H.run_parameters(ex)
    whole = K.philox_initial_occupancy(run.tables, 10, 6, seed=4)
    part = K.philox_initial_occupancy(run.tables, 4, 6, seed=4, traj_id0=6)
    assert np.array_equal(whole[6:], part)
    assert all(len(set(r)) == 6 for r in whole)
    assert np.all(run.tables.site_centre[whole] >= 0)


def test_library_builds_and_exports_every_declared_symbol():
    """nvcc cross-compiles for sm_100a without a GPU; every function declared in
    include/pycd_b200.h is exported (no compute calls here)."""
    path = nat.build()
    lib = ctypes.CDLL(str(path))
    header = (ROOT / 'include' / 'pycd_b200.h').read_text()
    declared = set(re.findall(r'\b(pycd_[a-z0-9_]+)\s*\(', header))
    assert declared == set(nat.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pycd_abi_version() == 3
    sass = subprocess.run(['cuobjdump', '-lelf', str(path)], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/pycd_b200.h against the ctypes table of pycd_b200/_native.py: same number of
    arguments (a mismatch would not fail at load time, it would corrupt a call)."""
    header = (ROOT / 'include' / 'pycd_b200.h').read_text()
    header = re.sub(r'/\*.*?\*/', ' ', header, flags=re.S)
    protos = re.findall(r'\b(pycd_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', header, flags=re.S)
    assert len(protos) == len(nat.EXPORTED_SYMBOLS)
    for name, args in protos:
        args = ' '.join(args.split())
        n = 0 if args in ('', 'void') else args.count(',') + 1
        assert name in nat._SIGNATURES, name
        assert len(nat._SIGNATURES[name][1]) == n, (name, n, len(nat._SIGNATURES[name][1]))


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """The four descriptor structs: size and the offset of every field as gcc lays them out from
    include/pycd_b200.h against the ctypes mirrors (same field names, same order)."""
    pairs = {'pycd_ewald_desc': nat.EwaldDesc, 'pycd_ewald_stats': nat.EwaldStats,
             'pycd_kmc_system_desc': nat.KmcSystemDesc, 'pycd_kmc_ensemble_desc': nat.KmcEnsembleDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "pycd_b200.h"}"',
             'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-o', str(exe), str(src)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f'{cname}.{fname}']) == getattr(cls, fname).offset, (cname, fname)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(nat.NativeError):
        nat.Context(0)


def test_msd_host_analysis_matches_reference_formulas():
    """analyse() (means, SEM, closed-form slope, D) vs the oracle's restatement of
    core.py:3028-3071 using scipy.stats.linregress like the reference."""
    import oracle as O
    from scipy.stats import linregress
    from pycd_b200 import msd as M
    rng = np.random.default_rng(3)
    n_traj, n_path, C_ = 3, 240, 4
    uw = np.cumsum(rng.normal(size=(n_traj, n_path, 3 * C_)), axis=1)
    uw[:, 0] = 0
    mp = M.MsdParameters(3, [4, 0], n_traj, (n_path - 0.5) * 1e-9, 1e-9, 100.5, 10, 300)
    assert mp.n_path == n_path and mp.n_msd == 101
    ref = O.msd_analysis(uw, [4, 0], mp.n_msd, mp.time_interval, mp.time_conversion, mp.dist_conversion,
                         mp.trim_length, 300, 3, linregress=linregress)
    res = M.analyse(mp, ref['species_avg_sd'])
    assert np.allclose(res['msd_data'], ref['msd_data'], rtol=1e-13, atol=0)
    assert np.allclose(res['sem_data'], ref['sem_data'], rtol=1e-13, atol=0)
    assert np.allclose(res['slopes'], ref['slopes'], rtol=1e-10)
    assert np.allclose(res['diffusivity'], ref['diffusivity'], rtol=1e-10)
    assert np.allclose(res['diffusivity_sem'], ref['diffusivity_sem'], rtol=1e-9)
    assert mp.file_tag == "1.00E+02ns_trim=10" or mp.file_tag == "1.01E+02ns_trim=10"
    with pytest.raises(ValueError):
        M.MsdParameters(3, [4, 0], 1, 1e-7, 1e-9, 5000.0, 10, 300)   # msd_t_final > t_final


def test_hdf5_writer_layout(tmp_path):
    h5py = pytest.importorskip('h5py')
    from pycd_b200.hdf5_io import write_trajectory_h5
    uw = np.arange(5 * 6, dtype=float).reshape(5, 6)
    assert write_trajectory_h5(tmp_path / 't.h5', uw, 2, 413.4137)
    with h5py.File(tmp_path / 't.h5') as f:
        assert f['coordinates'].shape == (5, 2, 3) and f['coordinates'].dtype == np.float32
        assert f['coordinates'].attrs['units'] == 'nanometers'
        assert f['time'].shape == (5,) and f['time'].attrs['units'] == 'picoseconds'
        assert list(f['topology/atoms/index'][:]) == [0, 1]


def test_hdf5_layout_through_the_h5py_api(tmp_path, monkeypatch):
    """The writer's LAYOUT (PyCD/hdf5_io.py:59-101, docs/hdf5_format.md:31-49) checked through an in-memory
    recorder of the h5py API (tests/fake_h5py.py): h5py is absent from this image and from the GPU box, so the
    real-file test above is skipped there; this one always runs.  Also: material_run's hdf5_output hook."""
    import importlib
    import fake_h5py
    monkeypatch.setitem(sys.modules, 'h5py', fake_h5py)
    import pycd_b200.hdf5_io as hio
    hio = importlib.reload(hio)
    try:
        assert hio.HDF5_AVAILABLE
        n_frames, n_car, dt_au = 7, 3, 413.4137
        rng = np.random.default_rng(0)
        uw = rng.normal(size=(n_frames, 3 * n_car)) * 40.0                      # bohr, (n_path, 3C) like unwrapped_traj.npy
        assert hio.write_trajectory_h5(tmp_path / 'trajectory.h5', uw, n_car, dt_au)
        f = fake_h5py.FILES[str(tmp_path / 'trajectory.h5')]
        assert set(f.keys()) == {'coordinates', 'time', 'topology'} and set(f['topology'].keys()) == {'atoms'}
        c, t = f['coordinates'], f['time']
        assert c.shape == (n_frames, n_car, 3) and c.dtype == np.float32 and c.compression == 'gzip' and c.chunks
        assert c.attrs['units'] == 'nanometers'
        # bohr -> nm: coords / ANG2BOHR * 0.1 (PyCD/hdf5_io.py:120, 157); 1 bohr = 0.0529177 nm
        assert np.allclose(c[:], (uw.reshape(n_frames, n_car, 3) * 0.052917721).astype(np.float32), rtol=1e-6)
        assert t.shape == (n_frames,) and t.dtype == np.float64 and t.compression == 'gzip'
        assert t.attrs['units'] == 'picoseconds'
        # a.u. of time -> ps: AUTIME2PS = 2.4188843e-5 (PyCD/hdf5_io.py:124, 158); frames sit on the grid tau * dt
        assert np.allclose(t[:], np.arange(n_frames) * dt_au * 2.418884326e-5, rtol=1e-8)
        atoms = f['topology/atoms']
        assert set(atoms.keys()) == {'name', 'element', 'index'}
        assert atoms['name'].dtype == np.dtype('S10') and list(atoms['name'][:]) == [b'ATOM0', b'ATOM1', b'ATOM2']
        assert atoms['element'].dtype == np.dtype('S2') and list(atoms['element'][:]) == [b'C'] * 3
        assert atoms['index'].dtype == np.int32 and list(atoms['index'][:]) == [0, 1, 2]
    finally:
        monkeypatch.delitem(sys.modules, 'h5py')
        importlib.reload(hio)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pycd_b200/ may mention it."""
    for f in (ROOT / 'pycd_b200').rglob('*'):
        if f.suffix in ('.py', '.cu', '.cuh', '.h'):
            src = f.read_text()
            assert 'oracle' not in src and 'ref_harness' not in src, f


def test_bench_cpu_legs_run_without_a_gpu():
    """bench.py's host-side pieces: the algorithmic-bytes formula of SURVEY 8(d), the Ewald CPU baseline
    (literal oracle on a sample of site pairs, extrapolated) and the reference arm's small fallback."""
    import json
    import subprocess
    from types import SimpleNamespace
    import bench
    assert bench.b_step_bytes(256, 64) == 275456          # cfg 3: 8*n_proc*(2C+6) + 4*n_proc
    assert bench.b_step_bytes(4, 1) == 272                # cfg 1
    lat, sc, run, ep = bench.build_problem(SimpleNamespace(size=[2, 2, 1], carriers=4))
    base = bench.ewald_cpu_baseline(sc, ep, n_sample=8)
    assert base['kind'] == 'port' and base['ns_per_pair_k_term'] > 0
    n = sc.num_system_elements
    assert np.isclose(base['seconds_full_array_extrapolated'],
                      base['ns_per_pair_k_term'] * 1e-9 * n * n * 3457, rtol=1e-12)
    out = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--cpu-seconds', '0.5'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'


def test_sweep_host_logic():
    """pycd_b200/sweep.py without a GPU: condition grid of BASELINE config 5, per-condition time grids that
    equalise the step counts (Arrhenius ratio of k_total), and the per-condition reduction -- trajectory g
    belongs to condition g % n_cond; diffusivity from the per-trajectory slopes (core.py:3052-3071), drift
    mobility from the accumulated drift (core.py:2052-2082)."""
    from pycd_b200 import sweep as SW
    from pycd_b200 import kmc as K
    conds = SW.hematite_conditions()
    assert len(conds) == 8 and [c[0] for c in conds] == [250.0, 250.0, 300.0, 300.0, 350.0, 350.0, 400.0, 400.0]
    assert not conds[0][1].any() and np.array_equal(conds[1][1], [1e-4, 0.0, 0.0])
    n_path, steps, C_ = 1001, 20000, 64
    iv = SW.balanced_intervals(conds, C_, steps, n_path)
    assert np.array_equal(iv[0::2], iv[1::2])                       # the field does not change the grid
    # at 300 K the grid spans steps / (C * 3.2e9 /s); the others follow the Arrhenius factor of the hop
    t300 = steps / (C_ * 3.2e9) * constants.SEC2AUTIME / (n_path - 1)
    assert np.isclose(iv[2], t300, rtol=1e-12)
    ratio = np.exp(-0.252 / 8.617333262e-5 * (1 / 250.0 - 1 / 300.0))
    assert np.isclose(iv[2] / iv[0], ratio, rtol=1e-12) and iv[0] > iv[2] > iv[4] > iv[6]
    # synthetic per-trajectory arrays: condition c has slope (c + 1) A^2 per ns, replica r adds r percent
    n_cond, reps, n_msd, trim = len(conds), 6, 41, 4
    n_total = n_cond * reps
    g = np.arange(n_total)
    slope = (g % n_cond + 1.0) * (1.0 + 0.01 * (g // n_cond))
    avg = np.empty((n_total, n_msd, 1))
    for t in range(n_total):
        dt_ns = iv[t % n_cond] * constants.AUTIME2NS
        avg[t, :, 0] = slope[t] * np.arange(n_msd) * dt_ns
    drift = np.zeros((n_total, C_, 3))
    drift[:, :, 0] = (g % n_cond + 1.0)[:, None] * 1e-3
    rows = SW.analyse_conditions(conds, iv, avg, drift, np.full(n_total, float(steps)), n_msd, trim)
    assert [r['n_traj'] for r in rows] == [reps] * n_cond
    for c, row in enumerate(rows):
        T, f = conds[c]
        kbt = constants.KB * T / constants.EV2J
        sl = slope[g % n_cond == c]
        want = sl.mean() * constants.ANG2CM ** 2 * constants.SEC2NS / 6 / kbt
        assert np.isclose(row['D_cm2_per_Vs'], want, rtol=1e-9)
        assert np.isclose(row['D_sem'], sl.std() / np.sqrt(reps) * constants.ANG2CM ** 2 * constants.SEC2NS / 6 / kbt, rtol=1e-7)
        assert row['mean_steps'] == steps
        if f.any():
            mob = K.drift_mobility(drift[g % n_cond == c], f, float(np.linalg.norm(f))).mean(axis=1)
            assert np.isclose(row['drift_mobility_cm2_per_Vs'], mob.mean(), rtol=1e-12)
            assert row['drift_mobility_sem'] == 0.0 or row['drift_mobility_sem'] < 1e-12 * abs(mob.mean())
        else:
            assert 'drift_mobility_cm2_per_Vs' not in row
