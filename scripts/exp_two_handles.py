"""Experiment: device-resident steps (pyramid build + B&B + refinement of 256 queries) alternating over
1, 2 or 3 handles, so that the HBM-bound pyramid build of one step overlaps the latency-bound sweep of another."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
import bench
NQ = 256
batch = bench.make_batch(0, NQ)
ids = np.arange(NQ, dtype=np.int64)
def make():
    h = capi.Handle(0)
    h.set_refiner(10, 1e-4, 1e-4, 1e4)
    bb = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=h)
    det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
    scan = matchers.ScanData(batch.angles[0], batch.ranges[0])
    queries = [matchers.LoopDetectionQuery(scan, 0, tuple(batch.scan_poses[i]),
               matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), i), tuple(batch.map_poses[i]), i)
               for i, s in enumerate(batch.submaps)]
    return h, det.prepare(queries), det
for nh in (1, 2, 3):
    hs = [make() for _ in range(nh)]
    streams = [torch.cuda.ExternalStream(h.stream) for h, _, _ in hs]
    res = (capi.CsmResult * NQ)(); ref = (capi.CsmRefined * NQ)()
    infl = [0] * nh
    def step(k):
        h, arr, _ = hs[k % nh]
        h.drop_pyramids(ids); h.build_pyramids(ids, 6)
        h.loop_batch_enqueue(arr, NQ, 6, 0)
        infl[k % nh] += 1
        if infl[k % nh] == 3:
            h.loop_batch_finish_refined(NQ, res, ref); infl[k % nh] -= 1
    def drain():
        for k in range(nh):
            while infl[k]:
                hs[k][0].loop_batch_finish_refined(NQ, res, ref); infl[k] -= 1
    for k in range(6): step(k)
    drain(); torch.cuda.synchronize()
    K = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]: s.wait_event(e0)
    for k in range(K): step(k)
    for s in streams[1:]: streams[0].wait_stream(s)
    e1.record(streams[0]); e1.synchronize()
    drain()
    print("%d handle(s): %.3f ms per step, found %d" % (nh, e0.elapsed_time(e1) / K, sum(r.found for r in res)))
    for h, _, _ in hs: h.close()
