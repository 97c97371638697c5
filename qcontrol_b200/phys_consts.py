"""Unit constants of the reference (phys_base.py:11-15), digit for digit.

Energies are in 1/cm, time in s, length in Bohr radii; every derived scalar the
kernels receive (t_sc, eL, the kinetic prefactor) is built from these on the host.
"""
import numpy

HART_TO_CM = numpy.float64(219474.6313708)   # 1 / cm / hartree
CM_TO_ERG = numpy.float64(1.98644568e-16)    # erg * cm
DALT_TO_AU = numpy.float64(1822.888486)      # a.u. / D
RED_PLANCK_H = numpy.float64(1.054572e-27)   # erg * s
HZ_TO_CM = numpy.float64(3.33564095e-11)     # s / cm
