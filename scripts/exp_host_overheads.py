"""Experiment: host-side cost of the per-step calls of bench.py's device-resident leg (run under torchrun)."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = capi.Handle(local)
h.set_refiner(10, 1e-4, 1e-4, 1e4)
batch = bench.make_batch(rank)
NQ = 256; ids = np.arange(NQ, dtype=np.int64)
bb = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=h)
det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
scan = matchers.ScanData(batch.angles[0], batch.ranges[0])
queries = [matchers.LoopDetectionQuery(scan, 0, tuple(batch.scan_poses[i]), matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), i), tuple(batch.map_poses[i]), i) for i, s in enumerate(batch.submaps)]
arr = det.prepare(queries)
res = (capi.CsmResult * NQ)(); ref = (capi.CsmRefined * NQ)()
def t(fn, n=50):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    el = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return el
h.loop_batch_enqueue(arr, NQ, 6, 0); h.loop_batch_finish_refined(NQ, res, ref)
w = torch.zeros(1, dtype=torch.int64, device="cuda")
view = torch.as_tensor(bench.CudaArrayView(h.best_key_device_ptr(), 1, "<i8"), device="cuda")
out = {}
out["drop+build_pyramids"] = t(lambda: (h.drop_pyramids(ids), h.build_pyramids(ids, 6)))
def enq():
    h.loop_batch_enqueue(arr, NQ, 6, 0); h.loop_batch_finish_refined(NQ, res, ref)
out["enqueue+finish (incl. GPU wait)"] = t(enq, 20)
out["as_tensor(view)"] = t(lambda: torch.as_tensor(bench.CudaArrayView(h.best_key_device_ptr(), 1, "<i8"), device="cuda"))
out["copy_"] = t(lambda: w.copy_(view))
out["Event()+record"] = t(lambda: torch.cuda.Event().record())
if world > 1:
    out["all_reduce"] = t(lambda: dist.all_reduce(w, op=dist.ReduceOp.MAX))
    out["all_reduce async_op"] = t(lambda: dist.all_reduce(w, op=dist.ReduceOp.MAX, async_op=True))
if rank == 0: print({k: round(v, 1) for k, v in out.items()})
if world > 1: dist.destroy_process_group()
