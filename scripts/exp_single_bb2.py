import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
case = synth.case_for(synth.CFG1, 41000)
s = case.submap
h = capi.Handle(0)
bb = matchers.ScanMatcherBranchBound("bb", 5, *synth.CFG2["rng"], handle=h)
gm = matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y))
scan = matchers.ScanData(case.angles, case.ranges)
for _ in range(50): bb.optimize_pose(gm, scan, tuple(case.init_pose))
t0 = time.perf_counter()
for _ in range(300):
    bb.optimize_pose(gm, scan, tuple(case.init_pose))
print("python-level per match us", (time.perf_counter() - t0) / 300 * 1e6)
key = bb._ensure_map(gm)[0]
sensor = matchers.compound(tuple(case.init_pose), scan.relative_sensor_pose)
step = matchers.compute_search_step(gm.resolution, scan.ranges); win = matchers.search_window(bb.range, step)
t0 = time.perf_counter()
for _ in range(300):
    h.match_bb(key, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
print("match_bb alone us", (time.perf_counter() - t0) / 300 * 1e6)
