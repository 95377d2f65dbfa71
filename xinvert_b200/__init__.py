"""xinvert_b200 -- B200-native (sm_100a) SOR elliptic inverter behind the
xinvert API.  See DESIGN.md; the C-ABI is in include/xinv.h."""
__version__ = "0.1.0"

from ._lib import Context, XinvError, default_context, device_count, pinned_empty  # noqa: F401
from .solvers import (invert_general_2D, invert_standard_2D, invert_standard_3D,  # noqa: F401
                      invert_general_3D, invert_general_bih_2D, invert_standard_1D, invert_standard_2D_test,
                      solve_general_3D, solve_general_bih_2D, solve_standard_1D, solve_standard_2D_test,
                      solve_general_2D, solve_general_2D_rows, solve_standard_2D, solve_standard_2D_front, solve_standard_2D_rows,
                      solve_standard_3D,
                      solve_standard_3D_rows)
from .core import inv_general2D, inv_standard2D, inv_standard3D  # noqa: F401
from .apps import (cal_flow, default_iParams, default_mParams, invert_Eliassen,  # noqa: F401
                   invert_GillMatsuno, invert_omega, invert_Poisson, invert_Stommel)
from .xrshim import DataArray  # noqa: F401
