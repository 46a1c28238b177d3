"""pansim_b200: B200 (sm_100a) implementation of Pansim's per-generation
Wright-Fisher step and pairwise distance pass, behind a C ABI
(include/pansim_b200.h, libpansim_b200.so) with a thin ctypes host mirror of the
reference's `Population` API. No CPU fallback."""
from .params import Derived, Params, derive, fmt_f64, validate  # noqa: F401
from .population import Pansim, PansimError, make_config, standard_deviation  # noqa: F401
from .group import PansimGroup  # noqa: F401

__version__ = "0.1.0"
