"""Drop-in mirrors of the reference's scripts/functions helpers (same names, arguments and error behaviour),
backed by the CUDA library.  ``sys.path.insert(1, 'scripts'); import functions.fct_misc as fct_misc`` in the
reference becomes ``from proj_roadsurf_b200.functions import fct_misc``."""
from . import fct_misc, fct_rasters, fct_statistics  # noqa: F401
