"""Test infrastructure ONLY: loader for the unmodified reference (vpasumarthi/PyCD).

Imports the reference package from /root/reference through the four compatibility
shims of SURVEY.md Appendix B (matplotlib stubs, yaml Loader default, np.load
allow_pickle default, ragged np.asarray fallback).  Used only by
tests/golden/make_golden.py (fixture generation, in the build container) and by
the optional differential tests that run when /root/reference is present.
Nothing in pycd_b200/ may import this module.
"""
import os
import shutil
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path(os.environ.get('PYCD_REFERENCE_ROOT', '/root/reference'))


def reference_available():
    return (REFERENCE_ROOT / 'PyCD' / 'core.py').exists()


_core = None


def load_reference():
    """Returns the reference's PyCD.core module (shimmed)."""
    global _core
    if _core is not None:
        return _core
    import numpy as np
    import yaml
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.offsetbox',
                 'matplotlib.ticker'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib.pyplot'].switch_backend = lambda *a, **k: None
    sys.modules['matplotlib.offsetbox'].AnchoredText = object
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if not getattr(yaml.load, '_pycd_shim', False):
        _yl = yaml.load

        def _yaml_load(stream, Loader=None):
            return _yl(stream, Loader=Loader or yaml.FullLoader)
        _yaml_load._pycd_shim = True
        yaml.load = _yaml_load
    if not getattr(np.load, '_pycd_shim', False):
        _nl = np.load

        def _np_load(*a, **k):
            return _nl(*a, **{'allow_pickle': True, **k})
        _np_load._pycd_shim = True
        np.load = _np_load
    if not getattr(np.asarray, '_pycd_shim', False):
        _na = np.asarray

        def _asarray(a, *args, **kw):
            try:
                return _na(a, *args, **kw)
            except ValueError as e:
                if 'inhomogeneous' not in str(e):
                    raise
                out = np.empty(len(a), dtype=object)
                for i, x in enumerate(a):
                    out[i] = x
                return out
        _asarray._pycd_shim = True
        np.asarray = _asarray
    sys.path.insert(0, str(REFERENCE_ROOT))
    import PyCD.core as core
    from unittest.mock import MagicMock
    core.plt = MagicMock()
    core.AnchoredText = MagicMock()
    _core = core
    return core


def stage_example(name, dst, species_count=None, extra=None, with_log=True):
    """Writable copy of examples/<name> with the F2 config edit (hdf5 write: 0)."""
    import numpy as np
    import yaml
    core = load_reference()
    src = REFERENCE_ROOT / 'examples' / name
    dst = Path(dst)
    if dst.exists():
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    for p in dst.rglob('*'):
        os.chmod(p, 0o755 if p.is_dir() else 0o644)
    os.chmod(dst, 0o755)
    shutil.rmtree(dst / 'traj1', ignore_errors=True)
    cfg_path = dst / 'simulation_parameters.yml'
    cfg = yaml.safe_load(open(cfg_path))
    cfg['output_data']['hdf5_output']['write'] = 0
    cfg['output_data']['hdf5_output']['enabled'] = False
    if species_count is not None:
        cfg['species_count'] = list(species_count)
    if extra:
        def merge(a, b):
            for k, v in b.items():
                if isinstance(v, dict) and isinstance(a.get(k), dict):
                    merge(a[k], v)
                else:
                    a[k] = v
        merge(cfg, extra)
    yaml.safe_dump(cfg, open(cfg_path, 'w'))
    if with_log:
        from PyCD.material_setup import material_setup
        material_setup(dst / 'InputFiles', np.array(cfg['system_size']),
                       np.array(cfg['pbc']), 0, 0, 1, 0, 0)
    return dst
