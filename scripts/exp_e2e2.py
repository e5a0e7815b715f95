"""Experiment: end-to-end C++ Detect (block-sparse host buffers) for several chunkings."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, hostapi, synth
import bench

lib = capi.load()
ctx = hostapi.Context(0)
hdet = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
h = capi.Handle.from_pointer(hdet.handle(), 0)
batch = bench.make_batch(0)
N = 256
parts = [synth.dense_to_blocks(s.grid, 4) for s in batch.submaps]
counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
nb = int(counts.sum())
blk_ptr = lib.csm_alloc_pinned(nb * 512); idx_ptr = lib.csm_alloc_pinned(nb * 4)
np.ctypeslib.as_array((C.c_uint16 * (nb * 256)).from_address(blk_ptr))[:] = np.concatenate([p[0].reshape(-1) for p in parts])
np.ctypeslib.as_array((C.c_int32 * nb).from_address(idx_ptr))[:] = np.concatenate([p[1] for p in parts])
ids = np.arange(N, dtype=np.int64)
offx = np.array([s.off_x for s in batch.submaps]); offy = np.array([s.off_y for s in batch.submaps])
res = batch.submaps[0].res
mp = np.ascontiguousarray(batch.map_poses); sp = np.ascontiguousarray(batch.scan_poses)
ang = np.ascontiguousarray(batch.angles[0]); rng = np.ascontiguousarray(batch.ranges[0])
out = (hostapi.HostSummary * N)()
h.set_option("accumulate_best_key", 1)

def step():
    hdet.clear_cache()
    n, _ = hdet.detect(N, None, blk_ptr, idx_ptr, counts.ctypes.data, 4, 512, 512, res, offx, offy, ids, mp, sp, ang, rng, out)
    return n

CFGS = ((128, 64), (64, 64), (256, 64))
for chunk, up in CFGS:
    hdet.configure(chunk_size=chunk | (up << 16), coarse_covariance=False)
    for _ in range(3): step()
    h.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): n = step()
    h.synchronize()
    print("batch %3d upload group %3d: %.3f ms per Detect (found %d)" % (chunk, up, (time.perf_counter() - t0) / 20 * 1e3, n))

# timeline of one Detect (all streams), best config
TL = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 64)
hdet.configure(chunk_size=TL[0] | (TL[1] << 16), coarse_covariance=False)
for _ in range(3): step()
h.synchronize()
h.set_option("timing", 2)
t0 = time.perf_counter(); step(); t1 = time.perf_counter()
print("host wall of this Detect: %.3f ms" % ((t1 - t0) * 1e3))
for name, ms in h.timings():
    print("  %8.3f ms  %s" % (ms, name))
