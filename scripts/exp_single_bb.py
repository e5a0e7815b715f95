"""Experiment: the single-scan branch-and-bound match of bench.py (cfg2), flags and per-phase timings."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, hostapi, matchers, synth
case = synth.case_for(synth.CFG1, 41000)
s = case.submap
h = capi.Handle(0)
bb = matchers.ScanMatcherBranchBound("bb", 5, *synth.CFG2["rng"], handle=h)
gm = matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y))
scan = matchers.ScanData(case.angles, case.ranges)
r = bb.optimize_pose(gm, scan, tuple(case.init_pose)).result
print("flags", r.flags, "found", r.found, "processed", r.n_processed, r.n_ignored)
h.set_option("timing", 1)
for _ in range(3):
    r = bb.optimize_pose(gm, scan, tuple(case.init_pose)).result
print(h.timings())
h.set_option("timing", 0)
t0 = time.perf_counter()
for _ in range(200):
    bb.optimize_pose(gm, scan, tuple(case.init_pose))
print("python-level per match us", (time.perf_counter() - t0) / 200 * 1e6)
ctx = hostapi.Context(0)
ctx.set_device_epilogue(True)
blocks, index, _, _ = synth.dense_to_blocks(s.grid)
f = lambda: ctx.match_blocks("bb", blocks.copy(), index, 4, s.grid.shape, s.res, (s.off_x, s.off_y), case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"])
for _ in range(50): f()
t0 = time.perf_counter()
for _ in range(300): f()
print("C++ plugin per match us", (time.perf_counter() - t0) / 300 * 1e6)
hh = capi.Handle.from_pointer(ctx.handle(), 0)
hh.set_option("timing", 1); f(); print(hh.timings())
