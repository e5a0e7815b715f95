"""Experiment: host-side and device-side timeline of one chunked Detect through the C ABI."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
import torch
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
import bench

lib = capi.load()
h = capi.Handle(0)
ext = torch.cuda.ExternalStream(h.stream, device=torch.device("cuda", 0))
batch = bench.make_batch(0)
N = 256
parts = [synth.dense_to_blocks(s.grid, 4) for s in batch.submaps]
counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
first = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
nb = int(first[-1])
blk_ptr = lib.csm_alloc_pinned(nb * 512); idx_ptr = lib.csm_alloc_pinned(nb * 4)
np.ctypeslib.as_array((C.c_uint16 * (nb * 256)).from_address(blk_ptr))[:] = np.concatenate([p[0].reshape(-1) for p in parts])
np.ctypeslib.as_array((C.c_int32 * nb).from_address(idx_ptr))[:] = np.concatenate([p[1] for p in parts])
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

def run(nch, verbose):
    CH = N // nch
    arrs = [det.prepare(queries[c*CH:(c+1)*CH]) for c in range(nch)]
    cid = [np.ascontiguousarray(ids[c*CH:(c+1)*CH]) for c in range(nch)]
    cc = [np.ascontiguousarray(counts[c*CH:(c+1)*CH]) for c in range(nch)]
    cox = [np.ascontiguousarray(offx[c*CH:(c+1)*CH]) for c in range(nch)]
    coy = [np.ascontiguousarray(offy[c*CH:(c+1)*CH]) for c in range(nch)]
    stamps = []
    evs = []
    def mark(name):
        stamps.append((name, time.perf_counter()))
        e = torch.cuda.Event(enable_timing=True); e.record(ext); evs.append((name, e))
    h.synchronize(); torch.cuda.synchronize()
    mark("start")
    for c in range(nch):
        h.upload_grids_blocks_ptr(cid[c], blk_ptr + int(first[c*CH]) * 512, idx_ptr + int(first[c*CH]) * 4, cc[c], 4, 32, 32, res, cox[c], coy[c])
    mark("uploads enq")
    h.upload_scan(0, scan.angles, scan.ranges)
    for c in range(nch):
        h.build_pyramids(cid[c], 6); mark("build%d enq" % c)
        h.loop_batch_enqueue(arrs[c], CH, 6, c * CH); mark("bb%d enq" % c)
    for c in range(nch):
        h.loop_batch_finish(CH); stamps.append(("finish%d" % c, time.perf_counter()))
    torch.cuda.synchronize()
    t0 = stamps[0][1]
    if verbose:
        print("chunks=%d host: " % nch + "  ".join("%s=%.0f" % (n, (t - t0) * 1e6) for n, t in stamps))
        print("          dev:  " + "  ".join("%s=%.0f" % (n, evs[0][1].elapsed_time(e) * 1e3) for n, e in evs))
    return (stamps[-1][1] - t0) * 1e3

for nch in (1, 2, 4, 8):
    for _ in range(3): run(nch, False)
    ts = [run(nch, False) for _ in range(10)]
    run(nch, True)
    print("chunks=%d total %.3f ms (min %.3f)" % (nch, np.mean(ts), min(ts)))
