"""MDTraj-style HDF5 trajectory output (layout of the reference's PyCD/hdf5_io.py:59-101 and
docs/hdf5_format.md): /coordinates (n_frames, n_atoms, 3) f32 nm gzip, /time (n_frames,) f64
ps gzip, /topology/atoms/{name S10, element S2, index i32}.  h5py is optional, exactly like
the reference guards it (core.py:28-32)."""
import numpy as np

from . import constants

try:
    import h5py
    HDF5_AVAILABLE = True
except ImportError:  # pragma: no cover - h5py is not in this image
    h5py = None
    HDF5_AVAILABLE = False


def write_trajectory_h5(path, unwrapped, n_carriers, time_interval_au):
    """Frames are the rows of the time grid; times are tau*dt (the reference passes the
    per-step time list, whose length differs and is silently dropped)."""
    if not HDF5_AVAILABLE:
        print('Warning: h5py is not installed; HDF5 trajectory output skipped')
        return False
    n_frames = unwrapped.shape[0]
    nm = (unwrapped.reshape(n_frames, n_carriers, 3) / constants.ANG2BOHR * 0.1).astype(np.float32)
    ps = np.arange(n_frames) * time_interval_au * constants.AUTIME2PS
    with h5py.File(path, 'w') as f:
        c = f.create_dataset('coordinates', data=nm, compression='gzip', chunks=True)
        c.attrs['units'] = 'nanometers'
        t = f.create_dataset('time', data=ps, compression='gzip')
        t.attrs['units'] = 'picoseconds'
        atoms = f.create_group('topology').create_group('atoms')
        atoms.create_dataset('name', data=np.array([f'ATOM{i}' for i in range(n_carriers)], dtype='S10'))
        atoms.create_dataset('element', data=np.array(['C'] * n_carriers, dtype='S2'))
        atoms.create_dataset('index', data=np.arange(n_carriers, dtype=np.int32))
    return True
