"""Experiment / repro: C++ Detect from heap-allocated blocks at the bench's size, a few configurations."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, hostapi, synth
import bench
n_maps, lanes, threads, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
batch = bench.make_batch(0, n_maps)
hb = bench.HostBatch(batch, 0, n_maps, hostapi, synth)
ctx = hostapi.Context(0)
det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
ug = int(sys.argv[5]) if len(sys.argv) > 5 else 64
chunk = int(sys.argv[6]) if len(sys.argv) > 6 else 128
tail = int(sys.argv[7]) if len(sys.argv) > 7 else 0
det.configure(chunk_size=chunk | ((tail // 16) << 12) | (ug << 16), coarse_covariance=False, query_index_base=0)
det.use_device_refiner(10, 1e-4, 1e-4)
det.set_lanes(lanes)
det.set_gather_threads(threads)
if len(sys.argv) > 8:
    det.set_first_group_divisor(int(sys.argv[8]))
for r in range(reps):
    det.clear_cache()
    t0 = time.perf_counter()
    n, _ = hb.detect(det)
    print("rep", r, "found", n, "ms %.3f" % ((time.perf_counter() - t0) * 1e3), flush=True)
det.close(); ctx.close()
print("done")
