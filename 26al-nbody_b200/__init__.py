"""26al-nbody_b200 -- B200-native (sm_100a) drop-in for the hot path of jweatson/26al-nbody:
Hermite-4 block-timestep direct-summation gravity (behind `gravity.evolve_model`,
al26_nbody.py:833) and the fused 26Al/60Fe disc-enrichment pass (al26_nbody.py:878-1086).

The directory name is not a Python identifier; import it with
    importlib.import_module("26al-nbody_b200")      # or:  import al26_b200  (alias at the repo root)
Hand-written CUDA lives in csrc/ behind the C-ABI of include/al26_b200.h; there is no CPU fallback.
"""
from . import units  # noqa: F401
from ._lib import Al26Error, Context, Group, SO_PATH, HEADER_PATH, device_count, dist_unique_id, load  # noqa: F401
from .particles import Particles, Channel  # noqa: F401
from .gravity import GravityCore, B200Gravity  # noqa: F401
from .enrichment import EnrichCore, decay_fractions, NINV, ROWS, ROW  # noqa: F401
from . import ic  # noqa: F401
from . import dist  # noqa: F401
from . import stellar, driver, checkpoint  # noqa: F401
from .stellar import StellarStub, YieldTables  # noqa: F401
from .yields_io import Yields  # noqa: F401

__all__ = ["units", "Al26Error", "Context", "Group", "device_count", "Particles", "Channel", "GravityCore", "B200Gravity",
           "EnrichCore", "decay_fractions", "ic", "load", "dist_unique_id"]
