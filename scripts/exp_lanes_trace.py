import sys, os, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, hostapi, synth
import bench
lib = capi.load()
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
lanes, up = int(sys.argv[1]), int(sys.argv[2])
ctx = hostapi.Context(0)
hdet = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
hdet.configure(chunk_size=256 | (up << 16), coarse_covariance=False)
hdet.use_device_refiner(10, 1e-4, 1e-4)
hdet.set_lanes(lanes)
def step():
    hdet.clear_cache()
    return hdet.detect(N, None, blk_ptr, idx_ptr, counts.ctypes.data, 4, 512, 512, res, offx, offy, ids, mp, sp, ang, rng, out)[0]
for _ in range(6): step()
os.environ["CSM_HOST_TRACE"] = "1"
t0 = time.perf_counter(); step(); print("wall %.0f us" % ((time.perf_counter() - t0) * 1e6))
del os.environ["CSM_HOST_TRACE"]
h = capi.Handle.from_pointer(hdet.handle(), 0)
h.set_option("timing", 2)
step()
for name, ms in h.timings():
    print("  %8.3f ms  %s" % (ms, name))
