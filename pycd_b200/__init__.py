"""pycd_b200 -- B200-native (sm_100a) hot path of vpasumarthi/PyCD behind PyCD's own
entry points.  `from pycd_b200 import material_setup, material_run, material_msd`
replaces `from PyCD...` for the Ewald precompute, the KMC step loop and the MSD analysis."""
from .material_setup import material_setup
from .material_preprod import material_preprod
from .material_run import material_run
from .material_msd import material_msd

__all__ = ['material_setup', 'material_preprod', 'material_run', 'material_msd']
