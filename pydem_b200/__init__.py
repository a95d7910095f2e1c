"""pydem_b200 -- B200-native implementation of pyDEM's hot path.

D-infinity slope/aspect stencil, upstream-contributing-area (UCA) sweep and TWI as
hand-written CUDA for sm_100a behind a C ABI (include/pydem_b200.h), exposed through a
``DEMProcessor`` with the reference's ``calc_slopes_directions / calc_uca / calc_twi``
surface (reference pydem/dem_processing.py:98-258, 587, 682, 1647), so it can stand in for
``pydem.DEMProcessor`` under ``ProcessManager``.
"""
__version__ = "0.1.0"

from .dem_processing import DEMProcessor  # noqa: F401
from ._lib import set_sweep, get_sweep  # noqa: F401,E402
