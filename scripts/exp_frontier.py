"""Experiment: structure of the branch-and-bound frontier of a cfg3 batch: per height, how many nodes,
how they cluster into groups of 8 adjacent angles of one (query, x, y), and how a warp's 8 consecutive
list entries are spread."""
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
for level in (5, 4, 3, 2, 1):
    h.set_option("bb_stop_level", level)
    h.loop_batch(arr, NQ, 6, 0)
    nodes = h.node_list(level)
    q, t, x, y = nodes.T
    g = np.unique(np.stack([q, x, y, t >> 3], axis=1), axis=0)
    for gs in (8, 16, 32):
        gg = np.unique(np.stack([q, x, y, t // gs], axis=1), axis=0)
        print("level %d: %7d nodes, groups of %2d adjacent angles: %6d -> occupancy %.2f" %
              (level, len(nodes), gs, len(gg), len(nodes) / (gs * len(gg))))
    # a warp = 8 consecutive list entries: distinct (q), distinct (q, x, y), angle span
    n8 = (len(nodes) // 8) * 8
    w = nodes[:n8].reshape(-1, 8, 4)
    dq = np.array([len(np.unique(a[:, 0])) for a in w[:4000]])
    dxy = np.array([len(np.unique(a[:, [0, 2, 3]], axis=0)) for a in w[:4000]])
    print("         per warp of 8 list entries: %.2f queries, %.2f (q, x, y) cells" % (dq.mean(), dxy.mean()))
    # per query: nodes
    cnt = np.bincount(q, minlength=NQ)
    print("         nodes per query: max %d, median %d, queries with none %d" % (cnt.max(), int(np.median(cnt)), int((cnt == 0).sum())))
h.set_option("bb_stop_level", 0)
