import sys
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
h = capi.Handle(0)
case = synth.case_for(synth.CFG1, 123)
s = case.submap
gm = matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y))
scan = matchers.ScanData(case.angles, case.ranges)
gs = matchers.ScanMatcherGridSearch("gs", 0.4, 0.4, 0.06, 0.05, 0.05, 0.004, handle=h)
for mode in (1, 0):
    h.set_option("window_mode", mode)
    r = gs.optimize_pose(gm, scan, tuple(case.init_pose)).result
    print(mode, r.found, r.best_x, r.best_y, r.best_t, r.sum_value, r.n_known)
