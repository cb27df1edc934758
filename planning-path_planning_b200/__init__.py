"""B200-native hot path of the DyMu planner (ESA-PRL/planning-path_planning).

Contents (only what the path needs):
  csrc/         CUDA kernels for sm_100a + the C ABI of include/dymu_cuda.h  -> libdymu_cuda.so
  src/          host-side drop-in ``PathPlanning_lib::DyMuPathPlanner``        -> libdymu_b200.so
  capi/         flat C view of that class (include/dymu_planner_c.h)
  shim/         minimal stand-ins for the Rock base-types headers the class interface names
  cuda_api.py   ctypes binding of the device C ABI
  planner_api.py ctypes binding of the flat planner API (also drives the reference oracle)
  synthetic.py  seeded synthetic inputs of the benchmark shapes
  build.py      in-tree nvcc/g++ build

The directory name contains a hyphen; import it through ``dymu_b200.load()`` at the
repository root (module name ``planning_path_planning_b200``).
"""
import os

from . import build as build_tools
from . import cuda_api, planner_api, sharding, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(HERE, "libdymu_b200.so")
CUDA_SO = cuda_api.CUDA_SO

_PLANNER_LIB = None


def planner_lib():
    """PlannerLib bound to libdymu_b200.so (the drop-in DyMuPathPlanner).  No fallback."""
    global _PLANNER_LIB
    if _PLANNER_LIB is None:
        if not os.path.exists(HOST_SO):
            raise RuntimeError("libdymu_b200.so is not built: run "
                               "`python planning-path_planning_b200/build.py`")
        cuda_api.load_library()
        _PLANNER_LIB = planner_api.PlannerLib(HOST_SO)
    return _PLANNER_LIB


def DyMuPathPlanner(risk_distance, reconnect_distance, risk_ratio, approach):
    return planner_lib().DyMuPathPlanner(risk_distance, reconnect_distance, risk_ratio, approach)
