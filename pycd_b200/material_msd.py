"""Drop-in for PyCD/material_msd.py:13-70 + Analysis.compute_msd's file handling
(core.py:2974-3081)."""
from datetime import datetime

import numpy as np

from . import _native as nat
from . import msd
from .config import input_directory, load_material_parameters, load_simulation_parameters
from .fileio import generate_report
from .lattice import Lattice


def material_msd(dst_path):
    start_time = datetime.now()
    sim = load_simulation_parameters(dst_path)
    inp = input_directory(dst_path, sim, parallel_aware=False)  # material_msd.py:27-33
    lattice = Lattice(load_material_parameters(inp))
    mp = msd.MsdParameters(sim['n_dim'], sim['species_count'], sim['n_traj'], sim['t_final'],
                           sim['time_interval'], sim['msd_t_final'], sim['trim_length'], sim['temp'],
                           sim['repr_time'], sim['repr_dist'])
    name = sim['output_data']['unwrapped_traj']['file_name']
    unwrapped = np.stack([np.load(dst_path / f'traj{i + 1}' / name)[:mp.n_path]
                          for i in range(mp.n_traj)])  # core.py:2982-2995
    ctx = nat.default_context()
    avg = msd.species_avg_sd(ctx, unwrapped, mp.n_traj, mp.n_path, mp.total_species, mp.n_msd,
                             mp.dist_conversion, mp.type_offsets)
    res = msd.analyse(mp, avg)
    np.save(dst_path / f'MSD_Data_{mp.file_tag}.npy', res['msd_data'])
    species = [s for i, s in enumerate(lattice.species_types) if mp.species_count[i] != 0]
    lines = []
    for k, sp in enumerate(species):  # core.py:3064-3071
        lines.append(f"Estimated value of {sp} diffusivity is: {res['diffusivity'][k]:.3e} cm2/Vs\n")
        lines.append(f"Standard error of mean in {sp} diffusivity is: "
                     f"{res['diffusivity_sem'][k]:.3e} cm2/Vs\n")
    generate_report(start_time, dst_path, f'MSD_Analysis_{mp.file_tag}', 1, ''.join(lines))
    _plot(res, mp, species, sim.get('display_error_bars', 0), dst_path)
    return None


def _plot(res, mp, species, error_bars, dst_path):
    """generate_msd_plot (core.py:3083-3132); skipped when matplotlib is not installed."""
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
    except ImportError:
        return
    fig, ax = plt.subplots()
    data, trim = res['msd_data'], mp.trim_length
    for k, sp in enumerate(species):
        ax.plot(data[:, 0], data[:, k + 1], 'o', label=sp)
        if error_bars:
            ax.errorbar(data[:, 0], data[:, k + 1], yerr=res['sem_data'][:, k], fmt='none', capsize=3)
        x, y = data[trim:-trim, 0], data[trim:-trim, k + 1]
        slope = msd.least_squares_slope(x, y)
        ax.plot(x, y.mean() + slope * (x - x.mean()), 'r', label=sp + '-fitted')
    ax.set_xlabel(f'Time ({mp.repr_time})')
    ax.set_ylabel(f'MSD ({mp.repr_dist}^2)')
    ax.legend()
    fig.savefig(str(dst_path / f'MSD_Plot_{mp.file_tag}.png'))
    plt.close(fig)
