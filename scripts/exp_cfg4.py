"""Experiment: cfg4 grid search (1080 beams, 0.025 m grid, window 4 m x 4 m x 60 deg, 0.1 deg) --
TMA shared-memory tile kernel against the plain global-memory kernel."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth

h = capi.Handle(0)
c4 = synth.case_for(synth.CFG4, 44000)
s = c4.submap
gm = matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), 7)
scan = matchers.ScanData(c4.angles, c4.ranges)
gs = matchers.ScanMatcherGridSearch("gs", *synth.CFG4["rng"], *synth.CFG4["step"], handle=h)
res = {}
for mode, name in ((1, "global-memory kernel"), (0, "TMA tile kernel")):
    h.set_option("window_mode", mode)
    r = gs.optimize_pose(gm, scan, tuple(c4.init_pose)).result
    t0 = time.perf_counter()
    for _ in range(5):
        r = gs.optimize_pose(gm, scan, tuple(c4.init_pose)).result
    dt = (time.perf_counter() - t0) / 5
    h.set_option("timing", 1)
    gs.optimize_pose(gm, scan, tuple(c4.init_pose))
    print("   phases:", ", ".join("%s %.0f us" % (k, v * 1e3) for k, v in h.timings()))
    h.set_option("timing", 0)
    res[mode] = (r.found, r.best_x, r.best_y, r.best_t, r.sum_value, r.n_known, r.normalized_score)
    full = 161 * 161 * 601 * 1080
    print("%-22s %.2f ms/match  %.2f T gathers/s  %.0f GB/s of u16 payload  result %s" % (
        name, dt * 1e3, full / dt / 1e12, full * 2 / dt / 1e9, res[mode][:6]))
print("identical:", res[0] == res[1])
