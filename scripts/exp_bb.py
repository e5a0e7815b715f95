"""Experiment: frontier sizes of the level-synchronous B&B under different incumbents."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
import bench

h = capi.Handle(0)
batch = bench.make_batch(0)
bb = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=h)
det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
scan = matchers.ScanData(batch.angles[0], batch.ranges[0])
queries = [matchers.LoopDetectionQuery(scan, 0, tuple(batch.scan_poses[i]),
           matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), i), tuple(batch.map_poses[i]), i)
           for i, s in enumerate(batch.submaps)]
arr = det.prepare(queries)
def run(label):
    res = h.loop_batch(arr, 256, 6, 0)
    t0 = time.perf_counter()
    for _ in range(10):
        res = h.loop_batch(arr, 256, 6, 0)
    dt = (time.perf_counter() - t0) / 10
    c = h.frontier_counts()
    print("%-28s frontier per height %s  sum %d  processed %d  found %d  %.3f ms" % (
        label, c[:7], sum(c), sum(r.n_processed for r in res), sum(r.found for r in res), dt * 1e3))
    return res
h.set_option("bb_dive", 1); r = run("dive (beam 8)")
h.set_option("bb_dive", 0); r0 = run("no dive (threshold only)")
h.set_option("bb_dive", 1); r = run("dive again")
same = all((a.found, a.best_x, a.best_y, a.best_t, a.sum_value) == (b.found, b.best_x, b.best_y, b.best_t, b.sum_value) for a, b in zip(r, r0))
print("dive and no-dive results identical:", same)
pos = [x.n_processed for x in r if x.found]; neg = [x.n_processed for x in r if not x.found]
print("processed per found query: mean %.0f max %d ; per not-found: mean %.0f max %d" % (np.mean(pos), max(pos), np.mean(neg), max(neg)))
