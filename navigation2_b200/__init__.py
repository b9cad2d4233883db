"""navigation2_b200 — B200-native MPPI trajectory optimiser (drop-in hot path of nav2_mppi_controller).

The package holds only what the path needs: `csrc/` (CUDA kernels + the C ABI of
include/mppi_b200.h), the ctypes binding (`capi`), ROS-parameter-shaped configuration (`params`)
and the Python host mirror of mppi::Optimizer (`optimizer`).  A C++ host mirror with the
reference's class names lives in `host/`.  Nothing here imports the CPU oracle under oracle/.
"""
from . import capi, params  # noqa: F401
from .optimizer import (  # noqa: F401
    ControllerException, CudaUnavailable, CycleInput, CycleResult, NoValidControl, Optimizer)

__all__ = ["capi", "params", "Optimizer", "CycleInput", "CycleResult", "ControllerException",
           "NoValidControl", "CudaUnavailable"]
