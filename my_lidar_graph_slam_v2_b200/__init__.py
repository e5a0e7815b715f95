"""B200-native correlative scan matching / loop detection hot path of
my-lidar-graph-slam-v2, behind the reference's ScanMatcher / LoopDetector
plugin interface.

  csrc/         CUDA kernels (sm_100a) + the C ABI of include/csm_b200.h
  host/         C++ adapter classes mirroring the reference's plugin interface
  capi.py       ctypes binding of the C ABI
  matchers.py   Python mirror of the plugin interface (used by tests / bench)
  synth.py      seeded synthetic submaps and scans
  build.py      in-tree nvcc build of libcsm_b200.so

The CUDA library is the only implementation: nothing here falls back to a CPU path.
"""
from . import build, capi, matchers, synth  # noqa: F401

__all__ = ["build", "capi", "matchers", "synth"]
