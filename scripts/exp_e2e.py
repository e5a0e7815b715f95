"""Experiment: where does the end-to-end loop-detection step spend its time?"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
import bench

lib = capi.load()
h = capi.Handle(0)
batch = bench.make_batch(0)
N, R, Cc = 256, 512, 512
cells = R * Cc
host_ptr = lib.csm_alloc_pinned(N * cells * 2)
host = np.ctypeslib.as_array((C.c_uint16 * (N * cells)).from_address(host_ptr)).reshape(N, R, Cc)
for m, s in enumerate(batch.submaps): host[m] = s.grid
ids = np.arange(N, dtype=np.int64)
offx = np.array([s.off_x for s in batch.submaps]); offy = np.array([s.off_y for s in batch.submaps])
res = batch.submaps[0].res
bb = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=h)
det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
scan = matchers.ScanData(batch.angles[0], batch.ranges[0])
queries = [matchers.LoopDetectionQuery(scan, 0, tuple(batch.scan_poses[i]),
           matchers.GridMap(None, s.res, (s.off_x, s.off_y), i), tuple(batch.map_poses[i]), i)
           for i, s in enumerate(batch.submaps)]
det._cached_maps.update(range(N)); det._cached_scans[0] = scan
h.upload_scan(0, scan.angles, scan.ranges)

def T(fn, reps=10):
    fn(); h.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    h.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

for nch in (1, 2, 4, 8):
    CH = N // nch
    cid = [np.ascontiguousarray(ids[c*CH:(c+1)*CH]) for c in range(nch)]
    cptr = [(C.c_void_p * CH)(*[host_ptr + m * cells * 2 for m in range(c*CH, (c+1)*CH)]) for c in range(nch)]
    cox = [np.ascontiguousarray(offx[c*CH:(c+1)*CH]) for c in range(nch)]
    coy = [np.ascontiguousarray(offy[c*CH:(c+1)*CH]) for c in range(nch)]
    def up():
        for c in range(nch): h.upload_grids_ptr(cid[c], cptr[c], R, Cc, res, cox[c], coy[c])
    def pyr():
        for c in range(nch):
            h.drop_pyramids(cid[c]); h.build_pyramids(cid[c], 6)
    arrs = [det.prepare(queries[c*CH:(c+1)*CH]) for c in range(nch)]
    def prep():
        for c in range(nch): det.prepare(queries[c*CH:(c+1)*CH])
    def bbq():
        for c in range(nch):
            h.loop_batch_enqueue(arrs[c], CH, 6, c*CH); h.loop_batch_finish(CH)
    def full():
        up()
        for c in range(nch):
            a = det.prepare(queries[c*CH:(c+1)*CH])
            h.build_pyramids(cid[c], 6)
            h.loop_batch_enqueue(a, CH, 6, c*CH); h.loop_batch_finish(CH)
    print("chunks=%d  upload %.2f ms | pyramid %.2f | prepare(host) %.2f | bb+finish %.2f | full %.2f" % (
        nch, T(up), T(pyr), T(prep), T(bbq), T(full)))
