"""ctypes access to libcsm_host.so, the C++ mirror of the reference's plugin
interface (host/include/csm_host). Used by the parity tests to drive the C++
adapter classes end to end; applications link the C++ classes directly."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcsm_host.so")


class HostSummary(C.Structure):
    _fields_ = [
        ("found", C.c_int32), ("best_x", C.c_int32), ("best_y", C.c_int32), ("best_t", C.c_int32),
        ("sum_value", C.c_int64), ("n_known", C.c_int32), ("flags", C.c_int32),
        ("score", C.c_double), ("norm_cost", C.c_double), ("est_pose", C.c_double * 3),
        ("cov", C.c_double * 9),
    ]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise FileNotFoundError("%s not built: run python my_lidar_graph_slam_v2_b200/host/build_host.py" % LIB)
        lib = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        lib.csm_host_context_create.restype = C.c_void_p
        lib.csm_host_context_create.argtypes = [C.c_int]
        lib.csm_host_context_destroy.argtypes = [C.c_void_p]
        lib.csm_host_context_handle.restype = C.c_void_p
        lib.csm_host_context_handle.argtypes = [C.c_void_p]
        lib.csm_host_cost.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                      dp, dp, C.c_int, dp, C.c_double, dp, dp]
        lib.csm_host_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                       C.c_double, C.c_double, dp, dp, C.c_int, dp, dp, C.c_int, dp, dp,
                                       C.c_double, C.c_double, C.c_double, C.POINTER(HostSummary)]
        lib.csm_host_match_blocks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp,
                                              C.c_int, dp, dp, C.c_int, dp, dp, C.c_double, C.c_double,
                                              C.c_double, C.POINTER(HostSummary)]
        lib.csm_host_loop_detect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                             dp, dp, C.POINTER(C.c_int64), dp, dp, dp, dp, C.c_int, C.c_int,
                                             dp, C.c_double, C.c_double, C.c_double, C.POINTER(HostSummary)]
        lib.csm_host_loop_detect_kind.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                                  C.c_double, dp, dp, C.POINTER(C.c_int64), dp, dp, dp, dp,
                                                  C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double,
                                                  C.POINTER(HostSummary), C.c_char_p, C.c_int]
        lib.csm_host_set_detect_concurrency.argtypes = [C.c_int]
        lib.csm_host_refine.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        dp, dp, C.c_int, dp, dp, C.c_int, C.c_double, dp, C.c_double,
                                        C.POINTER(HostSummary)]
        lib.csm_host_loopdet_use_linear_solver.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        lib.csm_host_loopdet_set_lanes.argtypes = [C.c_void_p, C.c_int]
        lib.csm_host_loopdet_best_word.argtypes = [C.c_void_p]
        lib.csm_host_loopdet_best_word.restype = C.c_uint64
        lib.csm_host_context_set_device_epilogue.argtypes = [C.c_void_p, C.c_int]
        lib.csm_host_context_set_device_final_matcher.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double,
                                                                  C.c_double]
        lib.csm_host_loopdet_use_device_refiner.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        lib.csm_host_loopdet_create.restype = C.c_void_p
        lib.csm_host_loopdet_create.argtypes = [C.c_void_p, C.c_int, dp, C.c_double, C.c_double, C.c_double]
        lib.csm_host_loopdet_destroy.argtypes = [C.c_void_p]
        lib.csm_host_loopdet_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.csm_host_loopdet_clear_cache.argtypes = [C.c_void_p]
        lib.csm_host_loopdet_handle.restype = C.c_void_p
        lib.csm_host_loopdet_handle.argtypes = [C.c_void_p]
        lib.csm_host_loopdet_detect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int, C.c_int, C.c_int, C.c_double, dp, dp,
                                                C.POINTER(C.c_int64), dp, dp, dp, dp, C.c_int,
                                                C.POINTER(HostSummary)]
        lib.csm_host_heap_maps_create.restype = C.c_void_p
        lib.csm_host_heap_maps_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.csm_host_heap_maps_destroy.argtypes = [C.c_void_p]
        lib.csm_host_loopdet_detect_heap.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                                     dp, dp, C.POINTER(C.c_int64), dp, dp, dp, dp, C.c_int,
                                                     C.POINTER(HostSummary)]
        lib.csm_host_loopdet_set_gather_threads.argtypes = [C.c_void_p, C.c_int]
        lib.csm_host_loopdet_capacity_retries.argtypes = [C.c_void_p]
        lib.csm_host_multidet_create.restype = C.c_void_p
        lib.csm_host_multidet_create.argtypes = [C.c_int, C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                                 C.c_int, C.c_double, C.c_double]
        lib.csm_host_multidet_destroy.argtypes = [C.c_void_p]
        lib.csm_host_multidet_use_nccl.argtypes = [C.c_void_p]
        lib.csm_host_multidet_clear_cache.argtypes = [C.c_void_p]
        lib.csm_host_multidet_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.csm_host_multidet_best_word.argtypes = [C.c_void_p]
        lib.csm_host_multidet_best_word.restype = C.c_uint64
        lib.csm_host_multidet_shard_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        lib.csm_host_multidet_detect.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_int, C.c_int, C.c_int, C.c_double, dp, dp,
                                                 C.POINTER(C.c_int64), dp, dp, dp, dp, C.c_int,
                                                 C.POINTER(HostSummary)]
        lib.csm_host_mapbuilder_create.restype = C.c_void_p
        lib.csm_host_mapbuilder_create.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                                   C.c_double, C.c_double]
        lib.csm_host_mapbuilder_destroy.argtypes = [C.c_void_p]
        lib.csm_host_mapbuilder_append.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, dp, C.c_double, C.c_double]
        lib.csm_host_mapbuilder_latest.argtypes = [C.c_void_p, dp, dp, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.csm_host_mapbuilder_match_rt.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, dp, C.c_int, dp, C.c_double,
                                                     C.POINTER(HostSummary)]
        lib.csm_host_slam_create.restype = C.c_void_p
        lib.csm_host_slam_create.argtypes = [C.c_void_p, dp, C.c_int]
        lib.csm_host_slam_destroy.argtypes = [C.c_void_p]
        lib.csm_host_slam_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int]
        lib.csm_host_slam_counters.argtypes = [C.c_void_p, dp]
        for name in ("num_scan_nodes", "num_local_maps", "num_edges", "num_loops"):
            getattr(lib, "csm_host_slam_" + name).argtypes = [C.c_void_p]
        for name in ("scan_nodes", "local_maps", "edges", "loops"):
            getattr(lib, "csm_host_slam_" + name).argtypes = [C.c_void_p, dp]
        lib.csm_host_slam_local_map_cells.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


SLAM_COUNTERS = ("scans_in", "scans_processed", "backend_steps", "backend_steps_with_candidates", "loop_queries",
                 "loops_detected", "optimizations", "degenerations", "optimizer_calls",
                 "t_latest_map", "t_match", "t_append", "t_backend", "t_detect")


class SlamPipeline:
    """C++ SlamPipeline (host/include/csm_host/slam_pipeline.hpp): the reference's front end and back end
    around the device matchers and device-resident maps. `settings`: slam_settings.pack(...)."""

    def __init__(self, ctx, settings):
        self.lib = load()
        self.ctx = ctx
        v, vp = _d(settings)
        self.p = self.lib.csm_host_slam_create(ctx.ctx, vp, len(v))
        assert self.p, "bad settings vector"

    def run(self, angles, ranges, odom_poses, time_stamps, min_range=0.01, max_range=50.0, finish=False):
        """ranges (n_scans, n_beams), odom_poses (n_scans, 3): ProcessScan per scan; returns the scans used"""
        a, ap = _d(angles)
        r, rp = _d(ranges)
        o, op = _d(odom_poses)
        t, tp = _d(time_stamps)
        assert r.ndim == 2 and r.shape[1] == len(a) and o.shape == (r.shape[0], 3) and len(t) == r.shape[0]
        return self.lib.csm_host_slam_run(self.p, r.shape[0], r.shape[1], ap, rp, op, tp, min_range, max_range,
                                          int(finish))

    def run_carmen(self, log, finish=True):
        """SlamPipeline::RunLog over the records of a CarmenLog; returns the scans used"""
        self.lib.csm_host_slam_run_carmen.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        return self.lib.csm_host_slam_run_carmen(self.p, log.p, int(finish))

    def record_metrics(self):
        self.lib.csm_host_slam_record_metrics.argtypes = [C.c_void_p]
        self.lib.csm_host_slam_record_metrics(self.p)

    def save_metrics(self, output_path):
        """writes output_path + '.metric.json' in the reference launcher's layout"""
        self.lib.csm_host_slam_save_metrics.argtypes = [C.c_void_p, C.c_char_p]
        rc = self.lib.csm_host_slam_save_metrics(self.p, output_path.encode())
        assert rc == 0, rc

    def counters(self):
        out = np.zeros(len(SLAM_COUNTERS))
        self.lib.csm_host_slam_counters(self.p, out.ctypes.data_as(C.POINTER(C.c_double)))
        return dict(zip(SLAM_COUNTERS, out.tolist()))

    def _table(self, what, width):
        n = getattr(self.lib, "csm_host_slam_num_" + {"scan_nodes": "scan_nodes", "local_maps": "local_maps",
                                                      "edges": "edges", "loops": "loops"}[what])(self.p)
        out = np.zeros((n, width))
        if n:
            getattr(self.lib, "csm_host_slam_" + what)(self.p, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def scan_nodes(self):
        """(n, 7): global pose, local pose, local map id"""
        return self._table("scan_nodes", 7)

    def local_maps(self):
        """(n, 10): global pose, first / last scan node, finished, rows, cols, offset x, y"""
        return self._table("local_maps", 10)

    def edges(self):
        """(n, 7): local map id, scan node id, inter-local-map, loop, relative pose"""
        return self._table("edges", 7)

    def loops(self):
        """(n, 6): local map id, scan node id, relative pose, normalized score"""
        return self._table("loops", 6)

    def local_map_cells(self, map_id, block_size=16):
        info = self.local_maps()[map_id]
        rows, cols = int(info[6]), int(info[7])
        dense = np.zeros((rows, cols), dtype=np.uint16)
        alloc = np.zeros((rows // block_size, cols // block_size), dtype=np.uint8)
        rc = self.lib.csm_host_slam_local_map_cells(self.p, map_id, dense.ctypes.data, dense.size, alloc.ctypes.data,
                                                    alloc.size)
        assert rc == 0, rc
        return dense, alloc

    def close(self):
        if self.p:
            self.lib.csm_host_slam_destroy(self.p)
            self.p = None


class MapBuilder:
    """C++ GridMapBuilderGPU: the reference's latest map (GridMapBuilder::UpdateLatestMap) built and kept on
    the device; append(pose, scan) per scan like the front end."""

    def __init__(self, ctx, resolution=0.05, patch_size=16, scans_for_latest_map=10, usable_range_min=0.01,
                 usable_range_max=50.0, prob_hit=0.62, prob_miss=0.46):
        self.lib = load()
        self.ctx = ctx
        self.p = self.lib.csm_host_mapbuilder_create(ctx.ctx, resolution, patch_size, scans_for_latest_map,
                                                     usable_range_min, usable_range_max, prob_hit, prob_miss)

    def set_guard_band(self, cells):
        self.lib.csm_host_mapbuilder_set_guard_band.argtypes = [C.c_double]
        self.lib.csm_host_mapbuilder_set_guard_band(float(cells))

    def last_exact_beams(self):
        self.lib.csm_host_mapbuilder_last_exact_beams.argtypes = [C.c_void_p]
        return self.lib.csm_host_mapbuilder_last_exact_beams(self.p)

    def set_fast_hit_points(self, on):
        self.lib.csm_host_mapbuilder_set_fast_hit_points.argtypes = [C.c_void_p, C.c_int]
        self.lib.csm_host_mapbuilder_set_fast_hit_points(self.p, int(on))

    def append(self, pose, angles, ranges, rel_pose=(0.0, 0.0, 0.0), min_range=0.01, max_range=50.0):
        a, ap = _d(angles)
        r, rp = _d(ranges)
        p, pp = _d(pose)
        q, qp = _d(rel_pose)
        return self.lib.csm_host_mapbuilder_append(self.p, pp, ap, rp, len(a), qp, min_range, max_range)

    def latest(self, cap_cells=1 << 22):
        """(dense u16 map, block allocation, (offset x, offset y), map pose, block size)"""
        geo = np.zeros(6)
        pose = np.zeros(3)
        dense = np.zeros(cap_cells, dtype=np.uint16)
        alloc = np.zeros(cap_cells // 64, dtype=np.uint8)
        dp = C.POINTER(C.c_double)
        rc = self.lib.csm_host_mapbuilder_latest(self.p, geo.ctypes.data_as(dp), pose.ctypes.data_as(dp),
                                                 dense.ctypes.data, cap_cells, alloc.ctypes.data, len(alloc))
        assert rc == 0
        rows, cols, bs = int(geo[0]), int(geo[1]), int(geo[2])
        return (dense[:rows * cols].reshape(rows, cols).copy(), alloc[:(rows // bs) * (cols // bs)].reshape(rows // bs, cols // bs).copy(),
                (geo[3], geo[4]), pose, bs)

    def match_rt(self, angles, ranges, init_pose, low_resolution, rng, rel_pose=(0.0, 0.0, 0.0), covariance_scale=1e4):
        a, ap = _d(angles)
        r, rp = _d(ranges)
        p, pp = _d(init_pose)
        q, qp = _d(rel_pose)
        g, gp = _d(rng)
        out = HostSummary()
        rc = self.lib.csm_host_mapbuilder_match_rt(self.p, ap, rp, len(a), qp, pp, low_resolution, gp,
                                                   covariance_scale, C.byref(out))
        assert rc == 0
        return out

    def close(self):
        if self.p:
            self.lib.csm_host_mapbuilder_destroy(self.p)
            self.p = None


class HeapMaps:
    """A batch of maps in the reference's storage: every allocated 16x16 block in its own heap allocation
    (grid_map.cpp:522-535). Built once from a contiguous block list; the loop detectors gather the blocks
    into page-locked staging themselves."""

    def __init__(self, blocks, index, counts, log2bs=4):
        self.lib = load()
        blocks = np.ascontiguousarray(blocks, dtype=np.uint16)
        index = np.ascontiguousarray(index, dtype=np.int32)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        self.n_maps = len(counts)
        self.p = self.lib.csm_host_heap_maps_create(blocks.ctypes.data, index.ctypes.data, counts.ctypes.data,
                                                    self.n_maps, log2bs)

    def close(self):
        if self.p:
            self.lib.csm_host_heap_maps_destroy(self.p)
            self.p = None


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def cost(grid, res, off, angles, ranges, sensor_pose, covariance_scale=1e4):
    lib = load()
    g = np.ascontiguousarray(grid, dtype=np.uint16)
    a, ap = _d(angles)
    r, rp = _d(ranges)
    p, pp = _d(sensor_pose)
    nc = C.c_double()
    cov = (C.c_double * 9)()
    lib.csm_host_cost(g.ctypes.data, g.shape[0], g.shape[1], res, off[0], off[1], ap, rp, len(a), pp,
                      covariance_scale, C.byref(nc), cov)
    return nc.value, np.array(cov)


def loop_search(scan_ids, scan_poses, map_ids, map_scan_min, map_scan_max, map_finished, accum_travel_dist,
                last_finished_scan_id, last_finished_map_id, travel_dist_threshold, node_dist_threshold,
                num_of_candidate_nodes):
    """LoopSearcherNearest::Search of the C++ mirror (CPU). Returns ([(query scan node, reference scan node,
    reference local map)], [squared node distance])."""
    lib = load()
    ip = C.POINTER(C.c_int)
    lib.csm_host_loop_search.argtypes = [C.c_int, ip, C.POINTER(C.c_double), C.c_int, ip, ip, ip, ip, C.c_double,
                                         C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, ip,
                                         C.POINTER(C.c_double), C.c_int]
    si = np.ascontiguousarray(scan_ids, dtype=np.int32)
    sp = np.ascontiguousarray(scan_poses, dtype=np.float64).reshape(-1)
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (map_ids, map_scan_min, map_scan_max, map_finished)]
    cap = max(int(num_of_candidate_nodes), 1)
    out = np.zeros(3 * cap, dtype=np.int32)
    dist = np.zeros(cap, dtype=np.float64)
    n = lib.csm_host_loop_search(len(si), si.ctypes.data_as(ip), sp.ctypes.data_as(C.POINTER(C.c_double)),
                                 len(arrs[0]), *[a.ctypes.data_as(ip) for a in arrs], accum_travel_dist,
                                 last_finished_scan_id, last_finished_map_id, travel_dist_threshold,
                                 node_dist_threshold, num_of_candidate_nodes, out.ctypes.data_as(ip),
                                 dist.ctypes.data_as(C.POINTER(C.c_double)), cap)
    return [tuple(int(v) for v in out[3 * i:3 * i + 3]) for i in range(n)], [float(d) for d in dist[:n]]


def refine(grid, res, off, angles, ranges, init_pose, rel_pose=(0.0, 0.0, 0.0), iterations_max=10,
           convergence_threshold=1e-4, lam=1e-4, covariance_scale=1e4):
    """ScanMatcherLinearSolver::OptimizePose (CPU, no device). Returns (summary, lambda after the call);
    summary.best_t holds the number of iterations."""
    lib = load()
    g = np.ascontiguousarray(grid, dtype=np.uint16)
    a, ap = _d(angles)
    r, rp = _d(ranges)
    p, pp = _d(init_pose)
    q, qp = _d(rel_pose)
    lam_c = C.c_double(lam)
    out = HostSummary()
    lib.csm_host_refine(g.ctypes.data, g.shape[0], g.shape[1], res, off[0], off[1], ap, rp, len(a), pp, qp,
                        iterations_max, convergence_threshold, C.byref(lam_c), covariance_scale, C.byref(out))
    return out, lam_c.value


def hill_climb(grid, res, off, angles, ranges, init_pose, rel_pose=(0.0, 0.0, 0.0), linear_step=0.1,
               angular_step=0.1, max_iterations=100, max_num_of_refinements=5, covariance_scale=1e4, greedy=None):
    """ScanMatcherHillClimbing::OptimizePose (CPU, no device) over the square-error cost, or over the
    greedy-endpoint cost when greedy = (MapResolution, HitAndMissedDist, OccupancyThreshold, KernelSize,
    ScalingFactor, StandardDeviation). summary.best_t holds the number of iterations, summary.best_x the
    number of step halvings."""
    lib = load()
    dp = C.POINTER(C.c_double)
    lib.csm_host_hill_climb.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, dp,
                                        C.c_int, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double,
                                        dp, C.POINTER(HostSummary)]
    g = np.ascontiguousarray(grid, dtype=np.uint16)
    a, ap = _d(angles)
    r, rp = _d(ranges)
    p, pp = _d(init_pose)
    q, qp = _d(rel_pose)
    out = HostSummary()
    gr = None
    if greedy is not None:
        gr_arr, gr = _d(greedy)
    lib.csm_host_hill_climb(g.ctypes.data, g.shape[0], g.shape[1], res, off[0], off[1], ap, rp, len(a), pp, qp,
                            linear_step, angular_step, max_iterations, max_num_of_refinements, covariance_scale,
                            gr, C.byref(out))
    return out


class Context:
    def __init__(self, device=0):
        self.lib = load()
        self.ctx = self.lib.csm_host_context_create(device)

    def handle(self):
        """The csm_handle (C ABI) this context runs on."""
        return self.lib.csm_host_context_handle(self.ctx)

    def set_device_epilogue(self, on=True):
        """Real-time correlative / branch-and-bound matches on this context take cost and covariance
        of the decided pose from the device instead of the CPU epilogue."""
        self.lib.csm_host_context_set_device_epilogue(self.ctx, int(on))

    def close(self):
        if self.ctx:
            self.lib.csm_host_context_destroy(self.ctx)
            self.ctx = None

    def set_device_final_matcher(self, iterations_max=10, convergence_threshold=1e-4, initial_lambda=1e-4,
                                 covariance_scale=1e4):
        """Real-time correlative / branch-and-bound matches on this context also run the reference's final
        matcher (ScanMatcherLinearSolver) on the pose they find, on the device: est_pose, norm_cost and cov of
        the summary are the refined ones. iterations_max <= 0 switches it off."""
        self.lib.csm_host_context_set_device_final_matcher(self.ctx, iterations_max, convergence_threshold,
                                                           initial_lambda, covariance_scale)

    def match(self, kind, grid, res, off, angles, ranges, init_pose, iparam, rng, step=(0, 0, 0),
              thr=(0.0, 0.0), rel_pose=(0.0, 0.0, 0.0), covariance_scale=1e4):
        g = np.ascontiguousarray(grid, dtype=np.uint16)
        a, ap = _d(angles)
        r, rp = _d(ranges)
        p, pp = _d(init_pose)
        q, qp = _d(rel_pose)
        rg, rgp = _d(rng)
        st, stp = _d(step)
        out = HostSummary()
        rc = self.lib.csm_host_match(self.ctx, {"rt": 0, "bb": 1, "grid": 2}[kind], g.ctypes.data,
                                     g.shape[0], g.shape[1], res, off[0], off[1], ap, rp, len(a), pp, qp,
                                     iparam, rgp, stp, thr[0], thr[1], covariance_scale, C.byref(out))
        assert rc == 0
        return out

    def match_blocks(self, kind, blocks, index, log2bs, shape, res, off, angles, ranges, init_pose, iparam, rng,
                     step=(0, 0, 0), thr=(0.0, 0.0), rel_pose=(0.0, 0.0, 0.0), covariance_scale=1e4):
        """match() with the map in block-sparse form: blocks (n, bs, bs) uint16, index (n,) int32."""
        b = np.ascontiguousarray(blocks, dtype=np.uint16)
        ix = np.ascontiguousarray(index, dtype=np.int32)
        a, ap = _d(angles)
        r, rp = _d(ranges)
        p, pp = _d(init_pose)
        q, qp = _d(rel_pose)
        rg, rgp = _d(rng)
        st, stp = _d(step)
        out = HostSummary()
        rc = self.lib.csm_host_match_blocks(self.ctx, {"rt": 0, "bb": 1, "grid": 2}[kind], b.ctypes.data,
                                            ix.ctypes.data, len(ix), log2bs, shape[0], shape[1], res, off[0],
                                            off[1], ap, rp, len(a), pp, qp, iparam, rgp, stp, thr[0], thr[1],
                                            covariance_scale, C.byref(out))
        assert rc == 0
        return out

    def loop_detect(self, grids, res, off_x, off_y, map_ids, map_poses, scan_poses, angles, ranges,
                    hmax, rng, thr, covariance_scale=1e4):
        g = np.ascontiguousarray(grids, dtype=np.uint16)
        nq = g.shape[0]
        ox, oxp = _d(off_x)
        oy, oyp = _d(off_y)
        ids = np.ascontiguousarray(map_ids, dtype=np.int64)
        mp, mpp = _d(np.asarray(map_poses).reshape(-1))
        sp, spp = _d(np.asarray(scan_poses).reshape(-1))
        a, ap = _d(angles)
        r, rp = _d(ranges)
        rg, rgp = _d(rng)
        out = (HostSummary * nq)()
        rc = self.lib.csm_host_loop_detect(self.ctx, nq, g.ctypes.data, g.shape[1], g.shape[2], res, oxp, oyp,
                                           ids.ctypes.data_as(C.POINTER(C.c_int64)), mpp, spp, ap, rp, len(a),
                                           hmax, rgp, thr[0], thr[1], covariance_scale, out)
        assert rc == 0
        return list(out)

    def loop_detect_kind(self, kind, grids, res, off_x, off_y, map_ids, map_poses, scan_poses, angles, ranges,
                         iparam, rng, step, thr, covariance_scale=1e4):
        """LoopDetector{Correlative, BranchBound, GridSearch} (kind 0 / 1 / 2) of the C++ mirror over one
        shared scan. Returns (per-query summaries, observed metrics {id: [values]})."""
        g = np.ascontiguousarray(grids, dtype=np.uint16)
        nq = g.shape[0]
        ox, oxp = _d(off_x)
        oy, oyp = _d(off_y)
        ids = np.ascontiguousarray(map_ids, dtype=np.int64)
        mp, mpp = _d(np.asarray(map_poses).reshape(-1))
        sp, spp = _d(np.asarray(scan_poses).reshape(-1))
        a, ap = _d(angles)
        r, rp = _d(ranges)
        rg, rgp = _d(rng)
        st, stp = _d(step if step is not None else (0.0, 0.0, 0.0))
        out = (HostSummary * nq)()
        buf = C.create_string_buffer(1 << 16)
        rc = self.lib.csm_host_loop_detect_kind(self.ctx, kind, nq, g.ctypes.data, g.shape[1], g.shape[2], res,
                                                oxp, oyp, ids.ctypes.data_as(C.POINTER(C.c_int64)), mpp, spp,
                                                ap, rp, len(a), iparam, rgp, stp, thr[0], thr[1],
                                                covariance_scale, out, buf, len(buf))
        assert rc == 0
        return list(out), _parse_metrics(buf.value.decode())


def map_update_table(odds, reference_table_end=True):
    """GridMapBuilderGPU::UpdateTable(odds): table[v] = the cell value after one update of a cell at v."""
    out = np.zeros(65536, dtype=np.uint16)
    lib = load()
    lib.csm_host_map_update_table.argtypes = [C.c_double, C.c_int, C.c_void_p]
    lib.csm_host_map_update_table(float(odds), int(reference_table_end), out.ctypes.data)
    return out


def set_detect_concurrency(n):
    """Matchers the Correlative / GridSearch loop detectors of Context.loop_detect_kind run at once."""
    load().csm_host_set_detect_concurrency(int(n))


def _parse_metrics(text):
    out = {}
    for item in text.split(";"):
        if "=" in item:
            k, v = item.split("=", 1)
            out[k] = [float(x) for x in v.split(",") if x]
    return out


class LoopDetector:
    """Persistent C++ LoopDetectorBranchBound (host/src/loop_detector.cpp) on one device context.
    detect() takes raw host pointers (dense `values` or block-sparse `blocks`/`index`/`counts`)."""

    def __init__(self, ctx, hmax, rng, thr, covariance_scale=1e4):
        self.lib = load()
        self.ctx = ctx
        rg, rgp = _d(rng)
        self.det = self.lib.csm_host_loopdet_create(ctx.ctx, hmax, rgp, thr[0], thr[1], covariance_scale)

    def use_linear_solver(self, iterations_max=10, convergence_threshold=1e-4, initial_lambda=1e-4,
                          covariance_scale=1e4):
        """Refine detected loops with the reference's default final matcher (CPU)."""
        self.lib.csm_host_loopdet_use_linear_solver(self.det, iterations_max, convergence_threshold,
                                                    initial_lambda, covariance_scale)

    def use_device_refiner(self, iterations_max=10, convergence_threshold=1e-4, initial_lambda=1e-4,
                           covariance_scale=1e4):
        """Refine detected loops with the same solver on the device, batched behind the search."""
        self.lib.csm_host_loopdet_use_device_refiner(self.det, iterations_max, convergence_threshold,
                                                     initial_lambda, covariance_scale)

    def set_lanes(self, n):
        """n pipeline lanes (n - 1 more device contexts on the same device): every upload group becomes
        its own search batch on the lane of its maps. Set after use_linear_solver / before detect."""
        self.lib.csm_host_loopdet_set_lanes(self.det, int(n))

    def best_word(self):
        """Packed best (key, global query index) word of the last detect() over all lanes."""
        return int(self.lib.csm_host_loopdet_best_word(self.det))

    def configure(self, chunk_size=128, coarse_covariance=True, query_index_base=0):
        self.lib.csm_host_loopdet_configure(self.det, chunk_size, int(coarse_covariance), query_index_base)

    def clear_cache(self):
        self.lib.csm_host_loopdet_clear_cache(self.det)

    def handle(self):
        """The csm_handle (C ABI) the detector runs on."""
        return self.lib.csm_host_loopdet_handle(self.det)

    def detect(self, nq, values, blocks, index, counts, log2bs, rows, cols, res, off_x, off_y, map_ids,
               map_poses, scan_poses, angles, ranges, out=None):
        out = out if out is not None else (HostSummary * nq)()
        n = self.lib.csm_host_loopdet_detect(
            self.det, nq, values, blocks, index, counts, log2bs, rows, cols, res,
            off_x.ctypes.data_as(C.POINTER(C.c_double)), off_y.ctypes.data_as(C.POINTER(C.c_double)),
            map_ids.ctypes.data_as(C.POINTER(C.c_int64)),
            map_poses.ctypes.data_as(C.POINTER(C.c_double)), scan_poses.ctypes.data_as(C.POINTER(C.c_double)),
            angles.ctypes.data_as(C.POINTER(C.c_double)), ranges.ctypes.data_as(C.POINTER(C.c_double)),
            len(angles), out)
        return n, out

    def detect_heap(self, nq, heap_maps, rows, cols, res, off_x, off_y, map_ids, map_poses, scan_poses,
                    angles, ranges, out=None):
        """detect() from maps whose blocks are separate heap allocations (HeapMaps): the detector gathers
        them into page-locked staging with its thread pool, group by group, behind the PCIe copies."""
        out = out if out is not None else (HostSummary * nq)()
        dp = C.POINTER(C.c_double)
        n = self.lib.csm_host_loopdet_detect_heap(
            self.det, nq, heap_maps.p, rows, cols, res, off_x.ctypes.data_as(dp), off_y.ctypes.data_as(dp),
            map_ids.ctypes.data_as(C.POINTER(C.c_int64)), map_poses.ctypes.data_as(dp),
            scan_poses.ctypes.data_as(dp), angles.ctypes.data_as(dp), ranges.ctypes.data_as(dp), len(angles), out)
        return n, out

    def set_first_group_divisor(self, n):
        self.lib.csm_host_loopdet_set_first_group_divisor.argtypes = [C.c_void_p, C.c_int]
        self.lib.csm_host_loopdet_set_first_group_divisor(self.det, int(n))

    def set_gather_threads(self, n):
        self.lib.csm_host_loopdet_set_gather_threads(self.det, int(n))

    def capacity_retries(self):
        return int(self.lib.csm_host_loopdet_capacity_retries(self.det))

    def close(self):
        if self.det:
            self.lib.csm_host_loopdet_destroy(self.det)
            self.det = None


class MultiGpuLoopDetector:
    """C++ LoopDetectorBranchBoundMultiGPU: one process, n_gpus devices, queries sharded by
    LocalMapId mod n_gpus, one host thread per GPU, results in query order."""

    def __init__(self, n_gpus, hmax, rng, thr, covariance_scale=1e4, lanes=1, refine=(10, 1e-4, 1e-4)):
        self.lib = load()
        self.n_gpus = n_gpus
        rg, rgp = _d(rng)
        it, conv, lam = refine if refine else (0, 0.0, 0.0)
        self.det = self.lib.csm_host_multidet_create(n_gpus, hmax, rgp, thr[0], thr[1], covariance_scale, lanes,
                                                     it, conv, lam)

    def use_nccl(self):
        self.lib.csm_host_multidet_use_nccl(self.det)

    def configure(self, chunk_size=128, upload_chunk=64, gather_threads=0):
        self.lib.csm_host_multidet_configure(self.det, chunk_size, upload_chunk, gather_threads)

    def clear_cache(self):
        self.lib.csm_host_multidet_clear_cache(self.det)

    def best_word(self):
        return int(self.lib.csm_host_multidet_best_word(self.det))

    def shard_sizes(self):
        out = (C.c_int * self.n_gpus)()
        self.lib.csm_host_multidet_shard_sizes(self.det, out)
        return list(out)

    def detect(self, nq, blocks, index, counts, heap_maps, log2bs, rows, cols, res, off_x, off_y, map_ids,
               map_poses, scan_poses, angles, ranges, out=None):
        out = out if out is not None else (HostSummary * nq)()
        dp = C.POINTER(C.c_double)
        n = self.lib.csm_host_multidet_detect(
            self.det, nq, blocks, index, counts, heap_maps.p if heap_maps is not None else None, log2bs, rows, cols,
            res, off_x.ctypes.data_as(dp), off_y.ctypes.data_as(dp), map_ids.ctypes.data_as(C.POINTER(C.c_int64)),
            map_poses.ctypes.data_as(dp), scan_poses.ctypes.data_as(dp), angles.ctypes.data_as(dp),
            ranges.ctypes.data_as(dp), len(angles), out)
        return n, out

    def close(self):
        if self.det:
            self.lib.csm_host_multidet_destroy(self.det)
            self.det = None


def carmen_records(lib, prefix, handle):
    """The records behind `handle` (csm_host_carmen_* here, the checker's orc_carmen_* in the tests: same
    layout) as a list of dicts: kind ('odom' / 'scan'), sensor_id, time_stamp, odom_pose, velocity, and for
    scans relative_sensor_pose, min/max range, min/max angle, angles, ranges."""
    dp = C.POINTER(C.c_double)
    f = lambda name: getattr(lib, prefix + name)
    f("count").argtypes = [C.c_void_p]
    f("total_beams").argtypes = [C.c_void_p]
    f("export").argtypes = [C.c_void_p, dp, dp, dp]
    f("sensor_id").argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    n = f("count")(handle)
    nb = f("total_beams")(handle)
    head = np.zeros((n, 15))
    angles = np.zeros(max(nb, 1))
    ranges = np.zeros(max(nb, 1))
    f("export")(handle, head.ctypes.data_as(dp), angles.ctypes.data_as(dp), ranges.ctypes.data_as(dp))
    out, at = [], 0
    buf = C.create_string_buffer(64)
    for i in range(n):
        f("sensor_id")(handle, i, buf, 64)
        h = head[i]
        rec = {"kind": "scan" if h[0] == 1.0 else "odom", "sensor_id": buf.value.decode(), "time_stamp": h[1],
               "odom_pose": h[2:5].copy(), "velocity": h[5:7].copy()}
        if rec["kind"] == "scan":
            k = int(h[14])
            rec.update(relative_sensor_pose=h[7:10].copy(), min_range=h[10], max_range=h[11], min_angle=h[12],
                       max_angle=h[13], angles=angles[at:at + k].copy(), ranges=ranges[at:at + k].copy())
            at += k
        out.append(rec)
    return out


class CarmenLog:
    """C++ CarmenLogReader (host/include/csm_host/carmen_log.hpp) over a log text or file."""

    def __init__(self, text=None, path=None):
        self.lib = load()
        self.lib.csm_host_carmen_load.restype = C.c_void_p
        self.lib.csm_host_carmen_load.argtypes = [C.c_char_p, C.c_int]
        self.lib.csm_host_carmen_destroy.argtypes = [C.c_void_p]
        self.p = self.lib.csm_host_carmen_load((path if path is not None else text).encode(), int(path is not None))
        if not self.p:
            raise FileNotFoundError(path)

    def records(self):
        return carmen_records(self.lib, "csm_host_carmen_", self.p)

    def close(self):
        if self.p:
            self.lib.csm_host_carmen_destroy(self.p)
            self.p = None


def write_carmen_log(path, ranges, odom_poses, time_stamps, start_angle, angular_resolution, max_range,
                     laser_on_robot=(0.0, 0.0, 0.0), old_format=False, with_odom=True):
    """A synthetic run as a Carmen log (CarmenLogWriter): ROBOTLASER1 records, or FLASER with PARAM records."""
    lib = load()
    dp = C.POINTER(C.c_double)
    lib.csm_host_carmen_write.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                          dp, dp, dp, dp, C.c_int]
    r, rp = _d(ranges)
    o, op = _d(odom_poses)
    t, tp = _d(time_stamps)
    l, lp = _d(laser_on_robot)
    rc = lib.csm_host_carmen_write(path.encode(), int(old_format), r.shape[0], r.shape[1], start_angle,
                                   angular_resolution, max_range, rp, op, lp, tp, int(with_odom))
    assert rc == 0, rc


def metric_values_string(metric_id, values):
    lib = load()
    lib.csm_host_metric_values_string.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_int]
    v, vp = _d(values)
    buf = C.create_string_buffer(64 * max(len(v), 1) + 16)
    lib.csm_host_metric_values_string(metric_id.encode(), vp, len(v), buf, len(buf))
    return buf.value.decode()


def metrics_json(sequences):
    """WriteMetricsJson for {id: values}; returns the file's text"""
    lib = load()
    lib.csm_host_metrics_json.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double),
                                          C.c_char_p, C.c_int]
    ids = list(sequences)
    counts = (C.c_int * max(len(ids), 1))(*[len(sequences[k]) for k in ids])
    flat = [float(x) for k in ids for x in sequences[k]]
    v, vp = _d(flat if flat else [0.0])
    buf = C.create_string_buffer(1 << 16)
    lib.csm_host_metrics_json("\n".join(ids).encode(), counts, len(ids), vp, buf, len(buf))
    return buf.value.decode()
