"""cfg4 grid search: the window kernels' CUDA-event times by window_mode (0 / 2: 32-bit TMA tiles, 3: u16 TMA tiles)."""
import sys
sys.path.insert(0, ".")
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
case = synth.case_for(synth.CFG4, 44000)
s = case.submap
h = capi.Handle(0)
gm = matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), 5)
scan = matchers.ScanData(case.angles, case.ranges)
mt = matchers.ScanMatcherGridSearch("gs", *synth.CFG4["rng"], *synth.CFG4["step"], handle=h)
full = 161 * 161 * 600 * 1080
for mode in (2, 3, 2, 3):
    h.set_option("window_mode", mode)
    mt.optimize_pose(gm, scan, tuple(case.init_pose))
    h.set_option("timing", 1)
    r = mt.optimize_pose(gm, scan, tuple(case.init_pose)).result
    t = dict(h.timings())
    h.set_option("timing", 0)
    ms = t.get("k_window_tma")
    print("mode", mode, "k_window_tma %.3f ms" % ms, "useful frac of 32*148*1.965e9: %.3f" % (full / (ms * 1e-3) / (32 * 148 * 1.965e9)),
          "best", (r.best_x, r.best_y, r.best_t), r.sum_value)
