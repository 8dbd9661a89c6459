"""B200-native per-step physics hot path of Simulation-Server (blood-cell simulation).

The product is ``libbcs.so`` (hand-written CUDA for sm_100a behind the C ABI of ``include/bcs.h``);
this package is the thin Python host layer over it (ctypes), plus scene/state helpers.
The directory name carries a hyphen, so import it with
``importlib.import_module("simulation-server_b200")``.
"""
from . import bcsd, scene, state  # noqa: F401
from .scene import CellDef, Scene, Layout, derive_layout, make_cylinder_vein, make_bifurcated_vein, make_rbc_celldef  # noqa: F401
from .state import make_initial_state  # noqa: F401
