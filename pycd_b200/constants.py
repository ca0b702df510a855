"""Unit conversions of the path (values identical to the reference's PyCD/constants.py:8-46).

The factors enter every number that crosses the drop-in boundary (bohr, Hartree,
atomic time), so they are derived from the same CODATA inputs in the same
operation order; tests/test_host.py pins the derived values.
"""
import math as _math

# SI inputs (PyCD/constants.py:8-19)
_EPS0 = 8.854187817E-12
_ANGSTROM = 1E-10
KB = 1.38064852E-23
_PI = 3.14159265358979323846
_M_E = 9.10938356E-31
_Q_E = 1.6021766208E-19
_HBAR = 1.054571800E-34

# derived atomic units (PyCD/constants.py:21-29)
_COULOMB_K = 1 / (4 * _PI * _EPS0)
BOHR = _HBAR ** 2 / (_M_E * _Q_E ** 2 * _COULOMB_K)
HARTREE = _HBAR ** 2 / (_M_E * BOHR ** 2)
AUTIME = _HBAR / HARTREE
_AU_TEMPERATURE = HARTREE / KB
_AU_POTENTIAL = HARTREE / _Q_E

# conversions (PyCD/constants.py:31-46)
EV2J = _Q_E
ANG2BOHR = _ANGSTROM / BOHR
ANG2UM = 1.00E-04
ANG2CM = 1.00e-08
J2HARTREE = 1 / HARTREE
SEC2AUTIME = 1 / AUTIME
V2AUPOT = 1 / _AU_POTENTIAL
EV2HARTREE = EV2J * J2HARTREE
SEC2NS = 1.00E+09
SEC2PS = 1.00E+12
SEC2FS = 1.00E+15
K2AUTEMP = 1 / _AU_TEMPERATURE
AUTIME2NS = SEC2NS / SEC2AUTIME
AUTIME2PS = SEC2PS / SEC2AUTIME
AUTIME2FS = SEC2FS / SEC2AUTIME
BOHR2UM = ANG2UM / ANG2BOHR
BOHR2CM = ANG2CM / ANG2BOHR

assert _math.isclose(ANG2BOHR, 1.8897261264678997, rel_tol=0, abs_tol=0)
