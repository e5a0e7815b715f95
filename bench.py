#!/usr/bin/env python
"""Benchmark of the B200-native loop-detection / scan-matching hot path.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ... # reference CPU path (oracle/_ref)

Workload (BASELINE.json configs[2]): one query scan (360 beams)
matched by branch-and-bound (hmax 6, window 2.5 m x 2.5 m x 0.5 rad, thresholds
0.55 / 0.6) against 256 candidate 512x512 submaps per GPU. A step is one
LoopDetector::Detect on 256 first-touch submaps: grid upload, pyramid build
(PrecomputeGridMaps), the batched B&B search and the refinement of every
detected loop by the reference's default final matcher (ScanMatcherLinearSolver,
10 iterations max, then its covariance), i.e. the reference's own cold path
(loop_detector_branch_bound.cpp:68-141). Queries shard across ranks
with no data-path exchange; the only collective is the 8-byte argmax
all-reduce of the packed best (score, query) word over NCCL.

  value : queries/s, inputs (submaps, scan) resident in HBM, CUDA events
  e2e   : queries/s through the C ABI with pinned HOST buffers, H2D of all
          submaps and D2H of all results inside the timed region
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HMAX = 6
N_MAPS = 256
ROWS = COLS = 512
REFINE = (10, 1e-4, 1e-4)      # NumOfIterationsMax, ConvergenceThreshold, InitialLambda (launcher_settings_default.json:28-35)
METRIC = "loop_detection_queries_per_sec"
UNIT = "queries/s"


emit = print


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_config(n_gpus):
    return {
        "workload": "cfg3: 1 query scan (360 beams) x %d candidate 512x512 u16 submaps per GPU, "
                    "branch-and-bound hmax=6, window 2.5m/2.5m/0.5rad, thr 0.55/0.6; step = Detect on "
                    "first-touch submaps (upload + pyramid build + batched B&B + linear-solver refinement "
                    "and covariance of every detected loop)" % N_MAPS,
        "queries_per_gpu": N_MAPS, "grid": "%dx%d u16 @0.05m" % (ROWS, COLS), "hmax": HMAX,
        "sharding": "queries/submaps sharded over %d rank(s), 8-byte NCCL argmax all-reduce" % n_gpus,
        "pipelining": "value: successive steps alternate over 2 handles per GPU (pyramid build of one step "
                      "overlaps the sweep of the previous one); e2e: 2 pipeline lanes inside one Detect",
        "l2": "inputs larger than L2: 128 MiB of submaps + 768 MiB of pyramid levels touched per step",
    }


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.active = False
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            if self.active:
                self.samples.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# data
# ---------------------------------------------------------------------------
def make_batch(rank, n_maps=N_MAPS):
    from my_lidar_graph_slam_v2_b200 import synth
    return synth.make_loop_batch(31000 + rank, n_maps=n_maps, true_fraction=0.25,
                                 rows=ROWS, cols=COLS, map_id_base=0)


def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_detect(kind, batch, n_threads, n_queries, cold=True, repeats=1):
    """Time LoopDetectorBranchBound::Detect on the CPU checker. Returns queries/s (best of repeats)."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import synth
    orc = pyoracle.load(kind)
    subs = batch.submaps[:n_queries]
    grids = [orc.grid(s.grid, s.res, s.off_x, s.off_y) for s in subs]
    det = orc.loop_detector(HMAX, synth.CFG3["rng"], synth.CFG3["thr"], n_threads)
    det.use_linear_solver(*REFINE)
    best = None
    for _ in range(repeats):
        if cold:
            det.clear_cache()
        _, el = det.detect(grids, batch.map_ids[:n_queries], batch.map_poses[:n_queries],
                           batch.scan_idx[:n_queries], batch.scan_poses[:n_queries],
                           batch.angles, batch.ranges)
        best = el if best is None else min(best, el)
    det.close()
    for g in grids:
        g.close()
    return n_queries / best


def reference_kind():
    from oracle import pyoracle
    return "reference" if pyoracle.available("reference") else "port"


# ---------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------
def main_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    kind = reference_kind()
    threads = cpu_threads()
    batch = make_batch(0)
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import synth
    orc = pyoracle.load(kind)
    grids = [orc.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    det = orc.loop_detector(HMAX, synth.CFG3["rng"], synth.CFG3["thr"], threads)
    det.use_linear_solver(*REFINE)

    def step(nq):
        det.clear_cache()
        _, el = det.detect(grids[:nq], batch.map_ids[:nq], batch.map_poses[:nq], batch.scan_idx[:nq],
                           batch.scan_poses[:nq], batch.angles, batch.ranges)
        return el

    # bounded sample: the first nq queries of the per-GPU step, nq chosen from one probe so
    # that the whole run stays within ~3 minutes whatever K the driver asks for
    nq = N_MAPS
    probe = step(min(N_MAPS, 4 * threads)) / min(N_MAPS, 4 * threads)
    budget_s = 150.0
    while nq > threads and probe * nq * (args.steps + args.warmup) > budget_s:
        nq //= 2
    times = []
    for it in range(args.warmup + args.steps):
        el = step(nq)
        if it >= args.warmup:
            times.append(el)
    total = float(np.sum(times))
    value = nq * args.steps / total
    sample = ("%d steps, each Detect on the first %d of the %d first-touch submaps of the per-GPU step "
              "(pyramid build + B&B + ScanMatcherLinearSolver refinement + covariance), queries split over %d std::threads"
              % (args.steps, nq, N_MAPS, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
class CudaArrayView:
    """Zero-copy torch view of a device word owned by the C library."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def main_cuda(args):
    import torch
    import torch.distributed as dist
    from my_lidar_graph_slam_v2_b200 import capi, hostapi, matchers, sharding, synth

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # the only collective is an 8-byte all-reduce that runs beside kernels filling every SM: one
        # channel of 64 threads is all it needs (a small CTA finds a free slot sooner), and NCCL's
        # stream gets high priority so that it is placed as soon as a slot frees
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        os.environ.setdefault("NCCL_NTHREADS", "64")
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
    lib = capi.load()
    # The reference-facing plugin: the C++ LoopDetectorBranchBound of host/ (libcsm_host.so) on top
    # of the C ABI. `h` is the csm_handle it runs on (device-side timing, best-word all-reduce).
    ctx = hostapi.Context(local)
    hdet = hostapi.LoopDetector(ctx, HMAX, synth.CFG3["rng"], synth.CFG3["thr"])
    # one search batch of 256 queries; first-touch submaps uploaded in 4 groups of 64 whose block
    # expansion + pyramid build overlap the PCIe transfer of the following groups
    # two pipeline lanes: search batches of 128 queries, first-touch submaps uploaded in groups of 64 on
    # one copy stream; the first batch is searched while the maps of the second still cross PCIe
    hdet.configure(chunk_size=128 | (64 << 16), coarse_covariance=False, query_index_base=rank * N_MAPS)
    # the reference's default final matcher on every detected loop, on the device (k_refine)
    hdet.use_device_refiner(*REFINE)
    hdet.set_lanes(2)
    h = capi.Handle.from_pointer(hdet.handle(), local)
    ext_stream = torch.cuda.ExternalStream(h.stream, device=torch.device("cuda", local))

    batch = make_batch(rank)
    cells = ROWS * COLS
    ids = np.arange(N_MAPS, dtype=np.int64)
    offx = np.array([s.off_x for s in batch.submaps])
    offy = np.array([s.off_y for s in batch.submaps])
    res = batch.submaps[0].res
    map_poses = np.ascontiguousarray(batch.map_poses, dtype=np.float64)
    scan_poses = np.ascontiguousarray(batch.scan_poses, dtype=np.float64)
    angles = np.ascontiguousarray(batch.angles[0], dtype=np.float64)
    ranges = np.ascontiguousarray(batch.ranges[0], dtype=np.float64)

    # Host inputs, page-locked. (a) the submaps in the reference's own storage form
    # (grid_map.cpp:262-266): only the 16x16 blocks that were ever written, back to back, plus
    # their positions -- what the adapter hands over without flattening; unallocated blocks never
    # cross PCIe. (b) the same submaps flattened to dense row-major u16 (comparison leg).
    LOG2BS = 4
    parts = [synth.dense_to_blocks(s.grid, LOG2BS) for s in batch.submaps]
    counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
    n_blocks = int(counts.sum())
    blk_bytes = 2 << (2 * LOG2BS)
    blk_ptr = lib.csm_alloc_pinned(max(n_blocks, 1) * blk_bytes)
    idx_ptr = lib.csm_alloc_pinned(max(n_blocks, 1) * 4)
    np.ctypeslib.as_array((C.c_uint16 * (n_blocks * (blk_bytes // 2))).from_address(blk_ptr))[:] = \
        np.concatenate([p[0].reshape(-1) for p in parts])
    np.ctypeslib.as_array((C.c_int32 * n_blocks).from_address(idx_ptr))[:] = np.concatenate([p[1] for p in parts])
    host_ptr = lib.csm_alloc_pinned(N_MAPS * cells * 2)
    host = np.ctypeslib.as_array((C.c_uint16 * (N_MAPS * cells)).from_address(host_ptr)).reshape(N_MAPS, ROWS, COLS)
    for m, s in enumerate(batch.submaps):
        host[m] = s.grid
    n_chunks = (N_MAPS + 63) // 64
    h2d_small = 2 * 360 * 8 + N_MAPS * (256 + 115 * 8 + 8 + 4)
    h2d_blocks = n_blocks * (blk_bytes + 4) + (N_MAPS + n_chunks) * 4
    summaries = (hostapi.HostSummary * N_MAPS)()
    best_word = torch.zeros(1, dtype=torch.int64, device="cuda")

    def allreduce_best(read_back=False):
        """8-byte all-reduce(max) of the packed best word over NCCL, on the handle's stream.
        read_back: also copy the result to the host (on that same stream) and return it."""
        view = torch.as_tensor(CudaArrayView(h.best_key_device_ptr(), 1, "<i8"), device="cuda")
        with torch.cuda.stream(ext_stream):
            best_word.copy_(view)
            sharding.allreduce_best(best_word)
            return int(best_word.item()) if read_back else None

    # Device-resident leg: the all-reduce of step k runs on a side stream behind an event, so that it
    # overlaps the kernels of step k + 1 instead of sitting between them (it is latency, not bandwidth)
    side_stream = torch.cuda.Stream(device=torch.device("cuda", local))
    word_ring = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(4)]
    word_done = [None] * 4
    ring_pos = [0]

    def allreduce_best_async(h=h, ext_stream=ext_stream):
        i = ring_pos[0] % 4
        ring_pos[0] += 1
        view = torch.as_tensor(CudaArrayView(h.best_key_device_ptr(), 1, "<i8"), device="cuda")
        with torch.cuda.stream(ext_stream):
            if word_done[i] is not None:
                ext_stream.wait_event(word_done[i])
            word_ring[i].copy_(view)                    # before the next batch clears the device word
            if world == 1:
                return
            ready = torch.cuda.Event()
            ready.record(ext_stream)
        with torch.cuda.stream(side_stream):
            side_stream.wait_event(ready)
            sharding.allreduce_best(word_ring[i])
            word_done[i] = torch.cuda.Event()
            word_done[i].record(side_stream)

    def e2e_step(sparse=True, hdet=hdet):
        """One LoopDetector::Detect of the C++ plugin on 256 first-touch submaps, from page-locked
        HOST buffers: upload (4 groups on the copy stream), block expansion, pyramid build, batched
        B&B, result read-back; then the 8-byte all-reduce of the best word and its read-back."""
        hdet.clear_cache()                                  # every submap is a first touch again
        h.set_option("reset_best_key", 1)
        n, _ = hdet.detect(N_MAPS, None if sparse else host_ptr, blk_ptr if sparse else None,
                           idx_ptr if sparse else None, counts.ctypes.data if sparse else None, LOG2BS,
                           ROWS, COLS, res, offx, offy, ids, map_poses, scan_poses, angles, ranges, summaries)
        # packed best (key, query) word of this rank over all lanes, reduced across ranks over NCCL
        word = hdet.best_word()
        if world > 1:
            with torch.cuda.stream(ext_stream):
                best_word.copy_(torch.tensor([word - (1 << 64) if word >= (1 << 63) else word], dtype=torch.int64))
                sharding.allreduce_best(best_word)
                word = int(best_word.item())
        return n, word

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- warm-up (also allocates every workspace) --------------------------------
    h.set_option("accumulate_best_key", 1)        # a Detect is several device batches
    for _ in range(max(args.warmup, 3)):
        e2e_step()

    # ---- e2e: host buffers, H2D + D2H inside the timed region ------------------------
    def time_e2e(sparse):
        for _ in range(2):
            e2e_step(sparse)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            n, w = e2e_step(sparse)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0), n, w

    # for comparison: the same Detect with the final matcher on the CPU (host/ ScanMatcherLinearSolver)
    e2e_cpu_refine_ms = None
    if world == 1:
        hdet_cpu = hostapi.LoopDetector(ctx, HMAX, synth.CFG3["rng"], synth.CFG3["thr"])
        hdet_cpu.use_linear_solver(*REFINE)
        hdet_cpu.configure(chunk_size=128 | (64 << 16), coarse_covariance=False, query_index_base=0)
        hdet_cpu.set_lanes(2)
        e2e_step(True, hdet_cpu)
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step(True, hdet_cpu)
        e2e_cpu_refine_ms = (time.perf_counter() - t0) / 3 * 1e3
        hdet_cpu.close()
        for _ in range(2):
            e2e_step()
    e2e_dense_s, n_dense, word_dense = time_e2e(False)
    sampler.active = True
    e2e_s, n_found, word = time_e2e(True)
    sampler.active = False
    assert (n_found, word) == (n_dense, word_dense), "block-sparse and dense uploads disagree"
    key, qidx = h.decode_best_key(word)

    # ---- value: inputs resident in HBM, CUDA events on the handles' streams ------------------
    # Successive steps (independent Detect calls) alternate over two handles, each with its own copy of
    # the 256 submaps resident: the HBM-bound pyramid build of one step overlaps the latency-bound
    # branch-and-bound sweep of the previous one. Every step does the full work of a step.
    h.set_option("accumulate_best_key", 0)
    scan = matchers.ScanData(angles, ranges)
    queries = [matchers.LoopDetectionQuery(
        scan, 0, tuple(batch.scan_poses[i]),
        matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), int(ids[i])), tuple(batch.map_poses[i]), i)
        for i, s in enumerate(batch.submaps)]
    N_LANES = 2
    lanes = []
    for k in range(N_LANES):
        hk = h if k == 0 else capi.Handle(local)
        hk.set_refiner(*REFINE)
        bb = matchers.ScanMatcherBranchBound("loop-bb", HMAX, *synth.CFG3["rng"], handle=hk)
        det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
        lanes.append({"h": hk, "det": det, "arr": det.prepare(queries),      # level 0 resident (untimed)
                      "stream": ext_stream if k == 0 else
                      torch.cuda.ExternalStream(hk.stream, device=torch.device("cuda", local)),
                      "in_flight": 0})
    arr = lanes[0]["arr"]
    results = (capi.CsmResult * N_MAPS)()
    refined = (capi.CsmRefined * N_MAPS)()
    for ln in lanes:
        ln["h"].synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    step_no = [0]

    def device_step():
        ln = lanes[step_no[0] % N_LANES]
        step_no[0] += 1
        hk = ln["h"]
        hk.drop_pyramids(ids)
        hk.build_pyramids(ids, HMAX)
        hk.loop_batch_enqueue(ln["arr"], N_MAPS, HMAX, rank * N_MAPS)   # includes the read-back of the results
        allreduce_best_async(hk, ln["stream"])
        ln["in_flight"] += 1
        if ln["in_flight"] == 3:                                # results of this lane's step two back
            hk.loop_batch_finish_refined(N_MAPS, results, refined)
            ln["in_flight"] -= 1

    def drain():
        for ln in lanes:
            while ln["in_flight"]:
                ln["h"].loop_batch_finish_refined(N_MAPS, results, refined)
                ln["in_flight"] -= 1

    for _ in range(3 * N_LANES):
        device_step()
    drain()
    barrier()
    launches0 = sum(ln["h"].launch_count() for ln in lanes)
    sampler.active = True
    ev[0].record(ext_stream)
    for ln in lanes[1:]:
        ln["stream"].wait_event(ev[0])               # no lane starts before the start mark
    t_issue = time.perf_counter()
    for _ in range(args.steps):
        device_step()
    host_issue_ms = (time.perf_counter() - t_issue) * 1e3 / args.steps
    for ln in lanes[1:]:
        ext_stream.wait_stream(ln["stream"])         # the end mark waits for every lane ...
    ext_stream.wait_stream(side_stream)              # ... and for the last all-reduces
    ev[1].record(ext_stream)
    ev[1].synchronize()
    sampler.active = False
    drain()
    launches = sum(ln["h"].launch_count() for ln in lanes) - launches0
    dev_ms = max_over_ranks(ev[0].elapsed_time(ev[1]))
    assert sum(r.found for r in results) == n_found == sum(f.valid for f in refined)
    for ln in lanes[1:]:
        ln["h"].close()

    # ---- per-kernel CUDA-event durations (library option "timing": one event after every kernel,
    # on the stream the kernels are launched on) for the roofline of the dominant kernels ----------
    barrier()
    h.set_option("timing", 1)
    kernel_ms = {}
    reps = max(5, min(args.steps, 20))
    for _ in range(reps):
        h.drop_pyramids(ids)
        h.build_pyramids(ids, HMAX)
        for k, v in h.timings():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v / reps
        h.loop_batch_enqueue(arr, N_MAPS, HMAX, rank * N_MAPS)
        h.loop_batch_finish(N_MAPS, results)
        for k, v in h.timings():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v / reps
    h.set_option("timing", 0)
    kernel_ms.pop("k_setup", None)          # its interval includes host staging time when the stream is idle
    counts = h.frontier_counts()
    pyr_ms = kernel_ms.get("k_pyramid_stream", 0.0)
    bb_ms = sum(v for k, v in kernel_ms.items() if k != "k_pyramid_stream")
    # nodes scored per step: the four children of every node of every list; the roots themselves are
    # expanded unscored (option "bb_skip_top"), counts[HMAX] = all root candidates
    nodes_scored = 4 * sum(counts[1:HMAX + 1])

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        pass
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            traffic = json.load(f)
    except (OSError, ValueError):
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    step_ms = dev_ms / args.steps
    # Dominant kernel family: k_bb_expand<HC> (one launch per pyramid height, 4 children per node).
    # Algorithmic bytes: one u16 grid read per scored node and beam.
    n_beams = len(angles)
    per_launch = {}
    for hc in range(HMAX):
        name = "k_bb_expand<%d>" % hc
        ms = kernel_ms.get(name, 0.0)
        nodes = 4 * counts[hc + 1]
        per_launch[name] = {"ms": ms, "nodes_scored": int(nodes),
                          "GBps": nodes * n_beams * 2 / (ms * 1e-3) / 1e9 if ms > 0 else None}
    exp_ms = sum(v["ms"] for v in per_launch.values())
    exp_bytes = sum(v["nodes_scored"] for v in per_launch.values()) * n_beams * 2
    roofline = {
        "kernel": "k_bb_expand<0..%d> (B&B frontier scoring, %d launches per step, one per pyramid height)"
                  % (HMAX - 1, HMAX),
        "bound": "hbm", "achieved": exp_bytes / (exp_ms * 1e-3) / 1e9 if exp_ms > 0 else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": exp_bytes / (exp_ms * 1e-3) / 1e9 / hbm_peak if exp_ms > 0 else None,
        "peak_source": peak_source, "algorithmic_bytes_per_step": int(exp_bytes),
        "avg_launch_ms": exp_ms / HMAX, "ms_per_step": exp_ms, "share_of_step": exp_ms / (pyr_ms + bb_ms),
        "traffic": traffic.get("k_bb_expand_dram_bytes_per_step"), "launches": per_launch,
        "note": "scattered 2-byte gathers from 900 MB of pyramid levels: neither DRAM nor tensor bound; the "
                "binding unit is the L1TEX line (wavefront) rate of divergent loads, see DESIGN.md section 5 "
                "and profiles/r1_kernels_full.txt, r1_bb_stalls.txt. Durations are CUDA events recorded by the library on the "
                "launching stream after every kernel.",
    }
    pyr_bytes = (1 + HMAX) * cells * 2 * N_MAPS          # read level 0 once, write hmax levels
    roofline_pyramid = {
        "kernel": "k_pyramid_stream (PrecomputeGridMaps, 1 launch per step)",
        "bound": "hbm", "achieved": pyr_bytes / (pyr_ms * 1e-3) / 1e9 if pyr_ms > 0 else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": pyr_bytes / (pyr_ms * 1e-3) / 1e9 / hbm_peak if pyr_ms > 0 else None, "peak_source": peak_source,
        "algorithmic_bytes_per_launch": pyr_bytes, "traffic": traffic.get("k_pyramid_stream_dram_bytes_per_launch"),
        "ms_per_step": pyr_ms, "share_of_step": pyr_ms / (pyr_ms + bb_ms),
    }
    phases = {"pyramid_ms": pyr_ms, "branch_and_bound_ms": bb_ms, "kernel_ms": kernel_ms,
              "note": "per-kernel durations of one step run alone on one handle (sum %.3f ms); the timed steps "
                      "alternate over two handles and overlap, so ms_per_step is below that sum" % (pyr_ms + bb_ms),
              "nodes_scored_per_step": int(nodes_scored)}

    total_queries = world * N_MAPS * args.steps
    line = {
        "metric": METRIC, "value": total_queries / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16/int64 (f64 projection)",
        "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": total_queries / e2e_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                "pipeline": "2 lanes (device contexts) x 128-query search batches, uploads in groups of 64",
                "h2d_bytes_per_step": h2d_blocks + h2d_small,
                "d2h_bytes_per_step": N_MAPS * (C.sizeof(capi.CsmResult) + C.sizeof(capi.CsmRefined)) + 16 + 8,
                "api": "C++ LoopDetectorBranchBound::Detect (host/, libcsm_host.so) over the C ABI, "
                       "final matcher = device refiner (csm_set_refiner)",
                "host_format": "block-sparse submaps (allocated 16x16 blocks + positions, the reference's "
                               "GridMap storage), %d of %d blocks allocated" % (n_blocks, N_MAPS * (ROWS >> 4) * (COLS >> 4))},
        "e2e_dense": {"value": total_queries / e2e_dense_s, "unit": UNIT,
                      "ms_per_step": 1e3 * e2e_dense_s / args.steps,
                      "h2d_bytes_per_step": N_MAPS * cells * 2 + h2d_small,
                      "host_format": "dense flattened submaps (csm_upload_grids)"},
        "e2e_cpu_final_matcher_ms_per_step": e2e_cpu_refine_ms,
        "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms,
        "roofline": roofline, "roofline_pyramid": roofline_pyramid, "phases": phases,
        "check": {"found_per_step": int(n_found), "best_key": int(key), "best_query": int(qidx)},
    }

    if rank == 0:
        sampler.stop()
        line["clocks"] = sampler.summary()
        if world == 1 and not args.no_cpu:
            kind = reference_kind()
            threads = cpu_threads()
            v_all = run_cpu_detect(kind, batch, threads, N_MAPS, cold=True, repeats=2)
            v_one = run_cpu_detect(kind, batch, 1, 32, cold=True)
            line["cpu_baseline"] = {
                "value": v_all, "unit": UNIT, "cores": threads, "kind": kind,
                "sample": "full step (256 queries on 256 first-touch submaps, refinement included), best of 2, %d threads; "
                          "1 thread on the first 32 queries: %.1f queries/s" % (threads, v_one),
                "one_core_value": v_one,
            }
            if not args.no_single:
                line["single_scan"] = single_scan_numbers(h, lib, kind)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    hdet.close()
    ctx.close()
    lib.csm_free_pinned(host_ptr)
    lib.csm_free_pinned(blk_ptr)
    lib.csm_free_pinned(idx_ptr)
    return 0


def single_scan_numbers(h, lib, kind):
    """matches/s of the three single-scan matchers through the C++ plugin classes of host/
    (ScanMatcher*::OptimizePose: host buffers in, grid upload + precompute + search + result
    read-back + CPU cost/covariance epilogue per match) next to the CPU checker on the same
    inputs (cfg1, cfg2, cfg4 of BASELINE.json)."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi, synth
    orc = pyoracle.load(kind)
    ctx = hostapi.Context(h.device)
    # cost / covariance of the decided pose come from the device, behind the match (csm_set_epilogue)
    ctx.set_device_epilogue(True)
    out = {"epilogue": "device (csm_set_epilogue) for real-time correlative and branch-and-bound, CPU for grid search"}

    def timeit(fn, reps, warm_s=0.5):
        # warm up for a fixed time, not a fixed count: these legs follow seconds of CPU-only work
        # during which the GPU drops to its idle clocks, and a 40 ms loop ends before they recover
        t0 = time.perf_counter()
        fn()
        while time.perf_counter() - t0 < warm_s:
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return reps / (time.perf_counter() - t0)

    case = synth.case_for(synth.CFG1, 41000)
    s = case.submap
    off = (s.off_x, s.off_y)
    og = orc.grid(s.grid, s.res, s.off_x, s.off_y)
    # the map in the reference's storage form (allocated 16x16 blocks); every match first copies the
    # blocks (what the adapter's gather out of the per-scan GridMap costs), then uploads them
    blocks, index, _, _ = synth.dense_to_blocks(s.grid)

    def rt():
        return ctx.match_blocks("rt", blocks.copy(), index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                                case.init_pose, 5, synth.CFG1["rng"])

    def bb():
        return ctx.match_blocks("bb", blocks.copy(), index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                                case.init_pose, 5, synth.CFG2["rng"])

    gpu = timeit(rt, 2000)
    cpu = timeit(lambda: orc.match_rt(og, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"]), 20, 0.0)
    out["cfg1_rt_matches_per_s"] = {"gpu_e2e": gpu, "cpu_1core": cpu, "ratio": gpu / cpu, "cpu_kind": kind}

    # the front end's pair: real-time correlative match, then the final matcher (linear solver) on the
    # pose found (lidar_graph_slam_frontend.cpp:216-230) -- on the device in one submission / on the CPU
    ctx_f = hostapi.Context(h.device)
    ctx_f.set_device_final_matcher(*REFINE)

    def rt_final():
        return ctx_f.match_blocks("rt", blocks.copy(), index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                                  case.init_pose, 5, synth.CFG1["rng"])

    def cpu_rt_final():
        o = orc.match_rt(og, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"])
        return orc.refine(og, case.angles, case.ranges, list(o.est_pose), None, *REFINE)

    gpu = timeit(rt_final, 2000)
    cpu = timeit(cpu_rt_final, 20, 0.0)
    out["cfg1_rt_plus_final_matcher_per_s"] = {"gpu_e2e": gpu, "cpu_1core": cpu, "ratio": gpu / cpu,
                                               "cpu_kind": kind}
    ctx_f.close()

    gpu = timeit(bb, 1000)
    cpu = timeit(lambda: orc.match_bb(og, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"]), 5, 0.0)
    out["cfg2_bb_matches_per_s"] = {"gpu_e2e": gpu, "cpu_1core": cpu, "ratio": gpu / cpu, "cpu_kind": kind}

    c4 = synth.case_for(synth.CFG4, 44000)
    s4 = c4.submap
    gpu = timeit(lambda: ctx.match("grid", s4.grid, s4.res, (s4.off_x, s4.off_y), c4.angles, c4.ranges,
                                   c4.init_pose, 0, synth.CFG4["rng"], step=synth.CFG4["step"]), 20, warm_s=0.1)
    # CPU: 1/64 of the window (x and y ranges / 8), scaled by the candidate ratio
    og4 = orc.grid(s4.grid, s4.res, s4.off_x, s4.off_y)
    rng_small = (synth.CFG4["rng"][0] / 8, synth.CFG4["rng"][1] / 8, synth.CFG4["rng"][2] / 8)
    t0 = time.perf_counter()
    r = orc.match_grid(og4, c4.angles, c4.ranges, c4.init_pose, rng_small, synth.CFG4["step"])
    el = time.perf_counter() - t0
    from my_lidar_graph_slam_v2_b200 import matchers as _m
    full = 1
    for r_, s_ in zip(synth.CFG4["rng"], synth.CFG4["step"]):
        full *= len(_m.grid_search_offsets(r_ / 2.0, s_))      # the reference's accumulating loops: 161 x 161 x 600
    cpu = 1.0 / (el * full / max(r.n_processed, 1))
    # the scoring kernel alone (CUDA events of the library) against the shared-memory roofline:
    # one LSU wavefront serves 32 candidate-beam reads, one wavefront per cycle per SM
    from my_lidar_graph_slam_v2_b200 import capi
    hh = capi.Handle.from_pointer(ctx.handle(), h.device)
    hh.set_option("timing", 1)
    ctx.match("grid", s4.grid, s4.res, (s4.off_x, s4.off_y), c4.angles, c4.ranges, c4.init_pose, 0,
              synth.CFG4["rng"], step=synth.CFG4["step"])
    km = dict(hh.timings())
    hh.set_option("timing", 0)
    wt_ms = km.get("k_window_tma")
    out["cfg4_grid_matches_per_s"] = {
        "gpu_e2e": gpu, "cpu_1core_scaled": cpu, "ratio": gpu / cpu,
        "cpu_sample": "%d of %d candidates evaluated in %.2f s, scaled linearly" % (r.n_processed, full, el),
        "cpu_kind": kind, "gather_bytes_per_match": full * 1080 * 2,
        "gather_GBps": full * 1080 * 2 * gpu / 1e9,
        "k_window_tma": None if not wt_ms else {
            "ms": wt_ms, "cell_reads_per_s": full * 1080 / (wt_ms * 1e-3),
            "smem_roofline_reads_per_s": 32 * 148 * 1.965e9,
            "frac_of_smem_roofline": full * 1080 / (wt_ms * 1e-3) / (32 * 148 * 1.965e9),
            "note": "roofline = 1 shared-memory wavefront (32 lanes) per cycle per SM at 1965 MHz; "
                    "ncu counts 0.739 wavefronts/cycle/SM including the tile widening and idle lanes (profiles/r1_k_window_tma.txt)"}}
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-single", action="store_true", help="skip the single-scan extras")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library prints there meanwhile (NCCL's
    # version banner, for one) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    global emit
    emit = lambda line: (out.write(line + "\n"), out.flush())
    if args.impl == "reference":
        return main_reference(args)
    return main_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
