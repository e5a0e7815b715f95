"""Experiment: end-to-end C++ Detect (block-sparse host buffers, device refiner) with pipeline lanes."""
import sys, time, ctypes as C
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
ref = None
for lanes, chunk, up in ((1, 256, 64), (2, 128, 64), (2, 128, 32), (4, 64, 64), (3, 96, 32), (2, 160, 32), (2, 192, 64), (4, 64, 32)):
    ctx = hostapi.Context(0)
    hdet = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
    hdet.configure(chunk_size=chunk | (up << 16), coarse_covariance=False)
    hdet.use_device_refiner(10, 1e-4, 1e-4)
    if lanes > 1:
        hdet.set_lanes(lanes)
    def step():
        hdet.clear_cache()
        n, _ = hdet.detect(N, None, blk_ptr, idx_ptr, counts.ctypes.data, 4, 512, 512, res, offx, offy, ids, mp, sp, ang, rng, out)
        return n
    for _ in range(4): step()
    t0 = time.perf_counter()
    for _ in range(30): n = step()
    el = (time.perf_counter() - t0) / 30 * 1e3
    sig = [(o.found, o.best_x, o.best_y, o.best_t, o.sum_value, tuple(o.est_pose)) for o in out]
    if ref is None: ref = (sig, hdet.best_word())
    same = sig == ref[0] and hdet.best_word() == ref[1]
    print("lanes %d batch %3d upload group %3d: %.3f ms per Detect (found %d, same results %s, best word %x)" % (lanes, chunk, up, el, n, same, hdet.best_word()))
    hdet.close(); ctx.close()
