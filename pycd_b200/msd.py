"""Host side of the MSD / diffusivity analysis (reference: Analysis, PyCD/core.py:2922-3081).
The lag loop (core.py:2996-3022) runs in csrc/msd.cu; means, standard errors and the
per-trajectory least-squares slopes are a few kB of numpy on the host."""
import numpy as np

from . import _native as nat
from . import constants


class MsdParameters:
    """Unit factors and step counts of Analysis.__init__ (core.py:2922-2972)."""

    def __init__(self, n_dim, species_count, n_traj, t_final, time_interval, msd_t_final,
                 trim_length, temp, repr_time='ns', repr_dist='angstrom'):
        self.n_dim = n_dim
        self.species_count = np.asarray(species_count, dtype=int)
        self.total_species = int(self.species_count.sum())
        self.n_traj = int(n_traj)
        self.t_final = t_final * constants.SEC2AUTIME
        self.time_interval = time_interval * constants.SEC2AUTIME
        self.trim_length = trim_length
        self.temp = temp
        self.n_path = int(self.t_final / self.time_interval) + 1
        self.repr_time, self.repr_dist = repr_time, repr_dist
        self.time_conversion = {'ns': constants.AUTIME2NS, 'ps': constants.AUTIME2PS,
                                'fs': constants.AUTIME2FS, 's': 1E+00 / constants.SEC2AUTIME}[repr_time]
        self.dist_conversion = {'m': constants.BOHR, 'um': constants.BOHR2UM,
                                'angstrom': 1E+00 / constants.ANG2BOHR}[repr_dist]
        self.kBT = constants.KB * self.temp / constants.EV2J  # eV
        self.msd_t_final = msd_t_final / self.time_conversion
        self.n_msd = int(self.msd_t_final / self.time_interval) + 1
        if self.n_msd > self.n_path:
            raise ValueError('msd_t_final exceeds t_final (the reference silently mis-indexes here)')
        if self.trim_length < 1:   # msd[trim:-trim] is EMPTY for trim = 0 (the reference's linregress raises on it)
            raise ValueError('trim_length must be >= 1: rows [trim:-trim] feed the diffusivity fit (core.py:3052-3058)')
        if 2 * self.trim_length >= self.n_msd:
            raise ValueError('trim_length leaves no points for the diffusivity fit')

    @property
    def type_offsets(self):
        counts = [int(c) for c in self.species_count if c != 0]
        return np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)

    @property
    def file_tag(self):  # core.py:3037-3041
        return ('%1.2E' % (self.msd_t_final * self.time_conversion)) + str(self.repr_time) \
            + '_trim=' + str(self.trim_length)


def species_avg_sd(ctx, unwrapped, n_traj, n_path, n_carriers, n_msd, scale, type_offsets,
                   want_carrier=False):
    """pycd_msd: (n_traj, n_msd, n_types) species-averaged squared displacement.  `unwrapped`
    may be a numpy array (n_traj, n_path, 3C), a torch CUDA tensor or a device address."""
    toff = np.ascontiguousarray(type_offsets, dtype=np.int32)
    n_types = len(toff) - 1
    out = np.empty((n_traj, n_msd, n_types))
    car = np.empty((n_traj, n_msd, n_carriers)) if want_carrier else None
    if isinstance(unwrapped, np.ndarray):
        unwrapped = np.ascontiguousarray(unwrapped, dtype=np.float64)
    nat.check(nat.lib().pycd_msd(ctx.handle, nat.ptr(unwrapped), int(n_traj), int(n_path),
                                 int(n_carriers), int(n_msd), float(scale), nat.ptr(toff), n_types,
                                 nat.ptr(out), nat.ptr(car)))
    return (out, car) if want_carrier else out


def least_squares_slope(x, y):
    """Slope of scipy.stats.linregress(x, y) (core.py:3056) in closed form."""
    xm = x.mean()
    return np.dot(x - xm, y - y.mean()) / np.dot(x - xm, x - xm)


def analyse(params, avg):
    """avg: (n_traj_total, n_msd, n_types) -> msd_data, sem_data, slopes, D, D_sem
    (core.py:3028-3071).  With several GPUs, `avg` is the all-gathered array."""
    n_traj = avg.shape[0]
    n_types = avg.shape[2]
    msd = np.zeros((params.n_msd, n_types + 1))
    msd[:, 0] = np.arange(params.n_msd) * params.time_interval * params.time_conversion
    msd[:, 1:] = np.mean(avg, axis=0)
    sem = np.std(avg, axis=0) / np.sqrt(n_traj)
    trim = params.trim_length
    x = msd[trim:-trim, 0]
    slopes = np.array([[least_squares_slope(x, avg[tr, trim:-trim, k]) for k in range(n_types)]
                       for tr in range(n_traj)])
    factor = constants.ANG2CM ** 2 * constants.SEC2NS / (2 * params.n_dim) / params.kBT
    diff = slopes.mean(axis=0) * factor
    diff_sem = slopes.std(axis=0) / np.sqrt(n_traj) * factor
    return {'msd_data': msd, 'sem_data': sem, 'slopes': slopes, 'diffusivity': diff,
            'diffusivity_sem': diff_sem}
