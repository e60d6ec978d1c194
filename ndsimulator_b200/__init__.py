"""ndsimulator_b200 -- B200-native multi-walker engine behind NDSimulator's YAML / NDRun interface.

The hot path (NDRun.run -> MD.update / Committor.update -> potential / colvar / bias evaluation) runs
as hand-written sm_100a CUDA (ndsimulator_b200/csrc) reached through the C ABI of include/ndsb200.h.
There is no CPU fallback: constructing an engine without the built library or without a GPU raises.
"""
from .config import parse as parse_config  # noqa: F401
from .engine import Engine, NdsError, map50to2d  # noqa: F401
from .ndrun import NDRun  # noqa: F401
from .plugin import B200Engine, register as register_reference_plugin  # noqa: F401  (method: b200 for the reference's NDRun)

__version__ = "0.1.0"
