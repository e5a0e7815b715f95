"""Experiment: per-kernel CUDA-event timings of a resident 256-query loop batch and the pyramid build."""
import sys
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
import bench

NQ = int(sys.argv[1]) if len(sys.argv) > 1 else 256
h = capi.Handle(0)
batch = bench.make_batch(0, NQ)
bb = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=h)
det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
scan = matchers.ScanData(batch.angles[0], batch.ranges[0])
queries = [matchers.LoopDetectionQuery(scan, 0, tuple(batch.scan_poses[i]),
           matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), i), tuple(batch.map_poses[i]), i)
           for i, s in enumerate(batch.submaps)]
arr = det.prepare(queries)
ids = np.arange(NQ, dtype=np.int64)
h.set_option("timing", 1)
if len(sys.argv) > 2 and sys.argv[2] == "refine":
    h.set_refiner(10, 1e-4, 1e-4, 1e4)       # the bench's step: final matcher on every found pose
import os
for kv in os.environ.get("CSM_OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); h.set_option(k, int(v))
acc = {}
REP = 20
for it in range(REP + 3):
    h.drop_pyramids(ids); h.build_pyramids(ids, 6)
    tp = h.timings()
    h.loop_batch(arr, NQ, 6, 0)
    tb = h.timings()
    if it >= 3:
        for k, v in tp + tb:
            acc[k] = acc.get(k, 0.0) + v
tot = 0.0
for k, v in acc.items():
    print("%-24s %8.1f us" % (k, v / REP * 1e3)); tot += v / REP * 1e3
print("%-24s %8.1f us   frontier %s" % ("sum", tot, h.frontier_counts()))
