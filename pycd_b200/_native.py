"""ctypes binding of libpycd_b200.so (the C ABI declared in include/pycd_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is
visible, the calls raise.  `build()` compiles the CUDA sources in-tree with nvcc for
sm_100a (it cross-compiles without a GPU).
"""
import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
CSRC_DIR = PKG_DIR / 'csrc'
LIB_PATH = PKG_DIR / 'libpycd_b200.so'
SOURCES = ['ctx.cu', 'ewald.cu', 'ewald_cells.cu', 'kmc.cu', 'kmc_stencil_nn4.cu', 'kmc_stencil_nn8.cu', 'kmc_stencil_nn12.cu',
           'msd.cu']
HEADERS = ['common.cuh', 'kmc_types.cuh', 'kmc_stencil.cuh', 'kmc_stencil_launch.h']
OBJ_DIR = CSRC_DIR / 'build'
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC']

KC_EWALD_FOURIER, KC_EWALD_FINISH, KC_EWALD_EXPAND, KC_KMC_STEP, KC_MSD, KC_VLAT = range(6)
RNG_REPLAY, RNG_PHILOX = 0, 1
P_DENSE, P_UNIT_ROWS = 0, 1


class NativeError(RuntimeError):
    pass


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).exists():
            return cand
    raise NativeError('nvcc not found; cannot build libpycd_b200.so')


def needs_build():
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC_DIR / s for s in SOURCES + HEADERS] + [PKG_DIR.parent / 'include' / 'pycd_b200.h']
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build(force=False, verbose=False, lib_path=None, extra=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> pycd_b200/libpycd_b200.so.
    The translation units compile in parallel (one nvcc -c each), then link.  lib_path / extra build a
    variant library beside the default one (A/B measurements: tools/)."""
    lib_path = Path(lib_path) if lib_path else LIB_PATH
    if not force and lib_path == LIB_PATH and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    extra = list(extra) if extra is not None else os.environ.get('PYCD_NVCC_EXTRA', '').split()  # e.g. -DPYCD_TRACE
    env = dict(os.environ)
    env.pop('CC', None)   # the image exports a gcc wrapper that nvcc must not pick up
    env.pop('CXX', None)
    obj_dir = OBJ_DIR / lib_path.stem
    obj_dir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = obj_dir / (Path(src).stem + '.o')
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', '-o', str(obj), str(CSRC_DIR / src)]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise NativeError(f'nvcc failed on {src}:\n' + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, err in results:
            print(err)
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(lib_path)] + \
          [str(o) for o, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise NativeError('nvcc link failed:\n' + res.stdout + res.stderr)
    return lib_path


class EwaldDesc(C.Structure):
    _fields_ = [('n_sites', C.c_int64), ('coords', C.c_void_p), ('cell', C.c_double * 9),
                ('pbc', C.c_int32 * 3), ('recip', C.c_double * 9), ('volume', C.c_double),
                ('alpha', C.c_double), ('r_cut', C.c_double), ('k_cut', C.c_double),
                ('dielectric', C.c_double), ('k_max', C.c_int32 * 3), ('k_part', C.c_int32),
                ('plan_rows', C.c_int64), ('k_parts', C.c_int32), ('reserved', C.c_int32)]


class EwaldStats(C.Structure):
    _fields_ = [('k_eff', C.c_int64), ('rows', C.c_int64), ('fourier_ms', C.c_double),
                ('finish_ms', C.c_double), ('flops', C.c_double), ('k_split', C.c_int32)]


class KmcSystemDesc(C.Structure):
    _fields_ = [('n_sites', C.c_int64), ('P', C.c_void_p), ('n_centres', C.c_int64),
                ('site_centre', C.c_void_p), ('site_class', C.c_void_p), ('n_class', C.c_int32),
                ('nn', C.c_int32), ('neigh', C.c_void_p), ('hopvec', C.c_void_p),
                ('lam', C.c_void_p), ('vab', C.c_void_p), ('e_rel', C.c_void_p),
                ('q_lat', C.c_void_p), ('q_carrier', C.c_double), ('kT', C.c_double),
                ('vn', C.c_double), ('field', C.c_double * 3), ('field_active', C.c_int32),
                ('p_layout', C.c_int32), ('n_basis', C.c_int32), ('size', C.c_int32 * 3)]


class KmcEnsembleDesc(C.Structure):
    _fields_ = [('n_traj', C.c_int64), ('traj_id0', C.c_uint64), ('n_carriers', C.c_int32),
                ('occupancy0', C.c_void_p), ('dt_grid', C.c_double), ('n_path', C.c_int64),
                ('step_limit', C.c_int64), ('stop_at_grid_end', C.c_int32),
                ('rng_mode', C.c_int32), ('seed', C.c_uint64), ('refresh_interval', C.c_int32),
                ('kT_traj', C.c_void_p), ('field_traj', C.c_void_p),
                ('record_unwrapped', C.c_int32), ('energy0', C.c_void_p),
                ('e_rel_traj', C.c_void_p), ('n_dopant_max', C.c_int32),
                ('dopant_site', C.c_void_p), ('dopant_dq', C.c_void_p), ('dt_grid_traj', C.c_void_p)]


_lib = None

_SIGNATURES = {
    'pycd_abi_version': (C.c_int, []),
    'pycd_last_error': (C.c_char_p, []),
    'pycd_ctx_create': (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    'pycd_ctx_destroy': (C.c_int, [C.c_void_p]),
    'pycd_ctx_info': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                C.POINTER(C.c_int64)]),
    'pycd_ctx_launch_count': (C.c_int64, [C.c_void_p]),
    'pycd_ctx_last_kernel_ms': (C.c_double, [C.c_void_p, C.c_int32]),
    'pycd_ctx_total_kernel_ms': (C.c_double, [C.c_void_p, C.c_int32]),
    'pycd_ctx_class_launches': (C.c_int64, [C.c_void_p, C.c_int32]),
    'pycd_ctx_reset_timers': (C.c_int, [C.c_void_p]),
    'pycd_ctx_flush_l2': (C.c_int, [C.c_void_p]),
    'pycd_host_alloc': (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    'pycd_host_free': (C.c_int, [C.c_void_p]),
    'pycd_ewald_rows': (C.c_int, [C.c_void_p, C.POINTER(EwaldDesc), C.c_int64, C.c_int64,
                                  C.c_void_p, C.POINTER(EwaldStats)]),
    'pycd_ewald_unit_rows': (C.c_int, [C.c_void_p, C.POINTER(EwaldDesc), C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                       C.POINTER(EwaldStats)]),
    'pycd_ewald_expand': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                                    C.c_int64, C.c_int64, C.c_void_p]),
    'pycd_kmc_system_create': (C.c_int, [C.c_void_p, C.POINTER(KmcSystemDesc),
                                         C.POINTER(C.c_void_p)]),
    'pycd_kmc_system_destroy': (C.c_int, [C.c_void_p]),
    'pycd_kmc_system_vlat': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pycd_kmc_ensemble_create': (C.c_int, [C.c_void_p, C.POINTER(KmcEnsembleDesc),
                                           C.POINTER(C.c_void_p)]),
    'pycd_kmc_ensemble_destroy': (C.c_int, [C.c_void_p]),
    'pycd_kmc_ensemble_reset': (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    'pycd_kmc_advance_async': (C.c_int, [C.c_void_p, C.c_int64]),
    'pycd_kmc_wait': (C.c_int, [C.c_void_p, C.c_void_p]),
    'pycd_kmc_advance': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(C.c_int64)]),
    'pycd_kmc_read': (C.c_int, [C.c_void_p] + [C.c_void_p] * 8),
    'pycd_kmc_read_energy': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'pycd_kmc_unwrapped_device': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    'pycd_kmc_last_kernel': (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    'pycd_nvtx_push': (C.c_int, [C.c_char_p]),
    'pycd_nvtx_pop': (C.c_int, []),
    'pycd_kmc_read_begin': (C.c_int, [C.c_void_p] + [C.c_void_p] * 7),
    'pycd_kmc_read_end': (C.c_int, [C.c_void_p]),
    'pycd_kmc_system_stencil': (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_char_p,
                                          C.c_int32]),
    'pycd_msd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64,
                           C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Loads (once) libpycd_b200.so; raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get('PYCD_B200_LIB') or LIB_PATH)   # PYCD_B200_LIB: a variant build (A/B measurements)
    if not path.exists():
        raise NativeError(f'{path} is missing: run `python -c "import __graft_entry__ as g; '
                          f'g.build()"` (there is no CPU fallback)')
    handle = C.CDLL(str(path))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise NativeError(lib().pycd_last_error().decode(errors='replace'))


def ptr(a):
    """void* of a numpy array, a torch tensor (host or device), an int address or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags['C_CONTIGUOUS']:
            raise ValueError('array must be C-contiguous')
        return a.ctypes.data
    if isinstance(a, int):
        return a
    if hasattr(a, 'data_ptr'):
        if not a.is_contiguous():
            raise ValueError('tensor must be contiguous')
        return a.data_ptr()
    raise TypeError(f'cannot take a pointer of {type(a)}')


def pinned_empty(shape, dtype=np.float64):
    """numpy array over page-locked host memory (kept alive for the life of the process)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().pycd_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Context:
    """One per process / GPU (pass LOCAL_RANK as device)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().pycd_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    @property
    def handle(self):
        if not self._h:
            raise NativeError('context already destroyed')
        return self._h

    def close(self):
        if self._h:
            lib().pycd_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        n_sm, free, total = C.c_int32(), C.c_int64(), C.c_int64()
        check(lib().pycd_ctx_info(self.handle, C.byref(n_sm), C.byref(free), C.byref(total)))
        return {'n_sm': n_sm.value, 'free_bytes': free.value, 'total_bytes': total.value}

    def launch_count(self):
        return lib().pycd_ctx_launch_count(self.handle)

    def last_kernel_ms(self, cls):
        return lib().pycd_ctx_last_kernel_ms(self.handle, cls)

    def total_kernel_ms(self, cls):
        return lib().pycd_ctx_total_kernel_ms(self.handle, cls)

    def class_launches(self, cls):
        return lib().pycd_ctx_class_launches(self.handle, cls)

    def reset_timers(self):
        check(lib().pycd_ctx_reset_timers(self.handle))

    def flush_l2(self):
        check(lib().pycd_ctx_flush_l2(self.handle))


class nvtx_range:
    """NVTX range around a phase of the host driver (Ewald / step batches / MSD), visible to nsys / ncu
    --nvtx.  Pushed through the library (nvtx3 is header-only inside libpycd_b200.so)."""

    def __init__(self, name):
        self.name = name.encode()

    def __enter__(self):
        lib().pycd_nvtx_push(self.name)
        return self

    def __exit__(self, *exc):
        lib().pycd_nvtx_pop()
        return False


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
