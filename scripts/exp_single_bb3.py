import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, hostapi, matchers, synth
case = synth.case_for(synth.CFG1, 41000)
s = case.submap
h = capi.Handle(0)
g = np.ascontiguousarray(s.grid)
def t(fn, n=300):
    for _ in range(30): fn()
    h.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    h.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print("upload_grid", t(lambda: h.upload_grid(5, g, s.res, s.off_x, s.off_y)))
print("upload+build_pyramid(5)", t(lambda: (h.upload_grid(5, g, s.res, s.off_x, s.off_y), h.build_pyramid(5, 5))))
print("upload+build_pyramid(5)+sync", t(lambda: (h.upload_grid(5, g, s.res, s.off_x, s.off_y), h.build_pyramid(5, 5), h.synchronize())))
bb = matchers.ScanMatcherBranchBound("bb", 5, *synth.CFG2["rng"], handle=h)
gm = matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y))
scan = matchers.ScanData(case.angles, case.ranges)
sensor = matchers.compound(tuple(case.init_pose), scan.relative_sensor_pose)
step = matchers.compute_search_step(gm.resolution, scan.ranges); win = matchers.search_window(bb.range, step)
def full():
    h.upload_grid(5, g, s.res, s.off_x, s.off_y); h.build_pyramid(5, 5)
    return h.match_bb(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
print("upload+build+match", t(full))
def nobuild():
    h.upload_grid(5, g, s.res, s.off_x, s.off_y)
    return h.match_bb(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
print("upload+match (levels on demand)", t(nobuild))
print("match alone", t(lambda: h.match_bb(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))))
h.set_option("timing", 2)
full()
print(h.timings())
