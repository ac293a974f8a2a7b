"""mcpm_b200 — thin Python (ctypes) harness over the C ABI of libmcpm.so.

The product is the CUDA library + C++ host driver under ../csrc; this package only exists so that the
tests and bench.py can drive the C ABI (include/mcpm.h) from Python.  It contains no numerics and no
fallback: if libmcpm.so is missing, or there is no CUDA device, calls raise.
"""
from .abi import (has_experiments, McpmError, SimConfig, SystemDesc, ChainStats, RunParams, ParticleEnergy, BoxEnergy, lib, lib_path,
                  RNG_PHILOX, RNG_REPLAY, GEOM_BULK, GEOM_SLIT, GEOM_CYLINDER, GEOM_SPHERE, WALL_NONE, WALL_SLIT_LJ, WALL_SPHERE_LJ,
                  PAIR_LJ, PAIR_HS)
from .engine import Engine, fp64_peak_tflops, device_count

__all__ = ["has_experiments", "McpmError", "SimConfig", "SystemDesc", "ChainStats", "RunParams", "ParticleEnergy", "BoxEnergy", "lib", "lib_path",
           "Engine", "fp64_peak_tflops", "device_count", "RNG_PHILOX", "RNG_REPLAY", "GEOM_BULK", "GEOM_SLIT", "GEOM_CYLINDER", "GEOM_SPHERE",
           "WALL_NONE", "WALL_SLIT_LJ", "WALL_SPHERE_LJ", "PAIR_LJ", "PAIR_HS"]
