"""Experiment: per-kernel timings of one real-time correlative match (cfg1)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
h = capi.Handle(0)
case = synth.case_for(synth.CFG1, 41000)
s = case.submap
scan = matchers.ScanData(case.angles, case.ranges)
sensor = matchers.compound(tuple(case.init_pose), scan.relative_sensor_pose)
step = matchers.compute_search_step(s.res, scan.ranges)
win = matchers.search_window(synth.CFG1["rng"], step)
h.upload_grid(5, np.ascontiguousarray(s.grid), s.res, s.off_x, s.off_y)
h.build_coarse(5, 5)
for epi in (0.0, 1e4):
    h.set_epilogue(epi)
    for _ in range(20): h.match_rt(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
    t0 = time.perf_counter()
    for _ in range(300): r = h.match_rt(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
    print("epilogue", epi, "match_rt %.1f us" % ((time.perf_counter() - t0) / 300 * 1e6))
    h.set_option("timing", 1)
    acc = {}
    for _ in range(20):
        h.match_rt(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
        for k, v in h.timings(): acc[k] = acc.get(k, 0) + v / 20
    h.set_option("timing", 0)
    print("   ", {k: round(v * 1e3, 1) for k, v in acc.items()})
