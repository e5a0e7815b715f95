"""ctypes loader for the two CPU checkers declared in oracle/oracle_api.h.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py. The product package
(my_lidar_graph_slam_v2_b200) never imports this module.

  load("reference") -> oracle/_ref/libcsm_ref.so  (unmodified reference TUs)
  load("port")      -> oracle/libcsm_port.so      (oracle/port.cpp restatement)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class OrcResult(C.Structure):
    _fields_ = [
        ("found", C.c_int32), ("best_x", C.c_int32), ("best_y", C.c_int32),
        ("best_t", C.c_int32), ("win_x", C.c_int32), ("win_y", C.c_int32),
        ("win_t", C.c_int32), ("n_known", C.c_int32),
        ("sum_value", C.c_int64),
        ("step_x", C.c_double), ("step_y", C.c_double), ("step_t", C.c_double),
        ("score", C.c_double), ("known_rate", C.c_double),
        ("n_processed", C.c_int32), ("n_ignored", C.c_int32),
        ("best_sensor_pose", C.c_double * 3), ("est_pose", C.c_double * 3),
        ("norm_cost", C.c_double), ("cov", C.c_double * 9),
    ]

    def asdict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


def build(kind=None):
    """Build the checkers with oracle/Makefile (ref only if /root/reference exists)."""
    targets = ["port", "ref", "adapter"] if kind is None else [{"reference": "ref"}.get(kind, kind)]
    subprocess.run(["make", "-s", "-C", HERE, "-j8"] + targets, check=True)


_PATHS = {
    "reference": os.path.join(HERE, "_ref", "libcsm_ref.so"),
    "port": os.path.join(HERE, "libcsm_port.so"),
    # the drop-in classes on the reference's own types + their driver (tests/integration), built with the
    # reference's translation units against the package's libcsm_b200.so
    "adapter": os.path.join(HERE, "_ref", "libcsm_adapter.so"),
}


def available(kind):
    return os.path.exists(_PATHS[kind])


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    def __init__(self, kind):
        path = _PATHS[kind]
        if not os.path.exists(path):
            build(kind)
        self.kind = kind
        lib = C.CDLL(path)
        self.lib = lib
        dp = C.POINTER(C.c_double)
        u16p = C.POINTER(C.c_uint16)
        i32p = C.POINTER(C.c_int32)
        rp = C.POINTER(OrcResult)
        lib.orc_kind.restype = C.c_char_p
        lib.orc_grid_create.restype = C.c_void_p
        lib.orc_grid_create.argtypes = [u16p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        lib.orc_grid_destroy.argtypes = [C.c_void_p]
        lib.orc_precompute.argtypes = [C.c_void_p, C.c_int, u16p]
        lib.orc_precompute_pyramid.argtypes = [C.c_void_p, C.c_int, u16p]
        common = [C.c_void_p, dp, dp, C.c_int, dp, dp]
        lib.orc_match_rt.argtypes = common + [C.c_int] + [C.c_double] * 5 + [rp]
        lib.orc_match_bb.argtypes = common + [C.c_int] + [C.c_double] * 5 + [rp]
        lib.orc_match_grid.argtypes = common + [C.c_double] * 8 + [rp]
        lib.orc_refine.argtypes = common + [C.c_int, C.c_double, dp, rp]
        lib.orc_loopdet_use_linear_solver.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        lib.orc_loopdet_create.restype = C.c_void_p
        lib.orc_loopdet_create.argtypes = [C.c_int] + [C.c_double] * 5 + [C.c_int]
        lib.orc_loopdet_destroy.argtypes = [C.c_void_p]
        lib.orc_loopdet_clear_cache.argtypes = [C.c_void_p]
        lib.orc_loopdet_detect.argtypes = [
            C.c_void_p, C.c_int, C.POINTER(C.c_void_p), i32p, dp, i32p, dp,
            C.c_int, C.c_int, dp, dp, rp, dp]
        if hasattr(lib, "orc_mapbuilder_create"):
            lib.orc_mapbuilder_create.restype = C.c_void_p
            lib.orc_mapbuilder_create.argtypes = [C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
            lib.orc_mapbuilder_append.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, dp, C.c_double, C.c_double]
            lib.orc_mapbuilder_latest.argtypes = [C.c_void_p, dp, dp, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        if hasattr(lib, "orc_slam_create"):
            lib.orc_slam_create.restype = C.c_void_p
            lib.orc_slam_create.argtypes = [dp, C.c_int]
            lib.orc_slam_destroy.argtypes = [C.c_void_p]
            lib.orc_slam_run.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int]
            lib.orc_slam_counters.argtypes = [C.c_void_p, dp]
            for name in ("num_scan_nodes", "num_local_maps", "num_edges", "num_loops"):
                getattr(lib, "orc_slam_" + name).argtypes = [C.c_void_p]
            for name in ("scan_nodes", "local_maps", "edges", "loops"):
                getattr(lib, "orc_slam_" + name).argtypes = [C.c_void_p, dp]
            lib.orc_slam_local_map_cells.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        assert lib.orc_kind().decode() == kind

    # -- grids ------------------------------------------------------------
    def grid(self, dense, res, off_x, off_y):
        dense = np.ascontiguousarray(dense, dtype=np.uint16)
        h = self.lib.orc_grid_create(
            dense.ctypes.data_as(C.POINTER(C.c_uint16)), dense.shape[0],
            dense.shape[1], res, off_x, off_y)
        if not h:
            raise ValueError("orc_grid_create failed (rows/cols must be multiples of 16)")
        return OracleGrid(self, h, dense.shape)

    # -- map construction (the reference's GridMapBuilder::UpdateLatestMap) ------------
    def map_builder(self, resolution=0.05, patch_size=16, scans_for_latest_map=10, usable_range_min=0.01,
                    usable_range_max=50.0, prob_hit=0.62, prob_miss=0.46):
        return OracleMapBuilder(self, resolution, patch_size, scans_for_latest_map, usable_range_min,
                                usable_range_max, prob_hit, prob_miss)

    def update_table(self, odds):
        """out[v] = cell value after GridBinaryBayes::UpdateOddsUnchecked(odds) of a cell at v (reference only)"""
        out = np.zeros(65536, dtype=np.uint16)
        self.lib.orc_update_table.argtypes = [C.c_double, C.c_void_p]
        assert self.lib.orc_update_table(float(odds), out.ctypes.data) == 0
        return out

    def value_probability(self, v):
        self.lib.orc_value_probability.restype = C.c_double
        self.lib.orc_value_probability.argtypes = [C.c_int]
        return self.lib.orc_value_probability(int(v))

    def slam(self, settings):
        """The full loop on the reference's own components (ref_wrapper.cpp: RefSlam); `settings` from
        my_lidar_graph_slam_v2_b200.slam_settings.pack()."""
        return OracleSlam(self, settings)

    # -- matchers ---------------------------------------------------------
    @staticmethod
    def _scan(angles, ranges, init_pose, rel_pose):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        p = np.ascontiguousarray(init_pose, dtype=np.float64)
        q = np.ascontiguousarray(rel_pose if rel_pose is not None else (0.0, 0.0, 0.0),
                                 dtype=np.float64)
        return a, r, p, q

    def match_rt(self, grid, angles, ranges, init_pose, low_res, rng, thr=(0.0, 0.0), rel_pose=None):
        a, r, p, q = self._scan(angles, ranges, init_pose, rel_pose)
        out = OrcResult()
        rc = self.lib.orc_match_rt(grid.h, _dptr(a), _dptr(r), len(a), _dptr(p), _dptr(q),
                                   low_res, rng[0], rng[1], rng[2], thr[0], thr[1], C.byref(out))
        assert rc == 0
        return out

    def match_bb(self, grid, angles, ranges, init_pose, hmax, rng, thr=(0.0, 0.0), rel_pose=None):
        a, r, p, q = self._scan(angles, ranges, init_pose, rel_pose)
        out = OrcResult()
        rc = self.lib.orc_match_bb(grid.h, _dptr(a), _dptr(r), len(a), _dptr(p), _dptr(q),
                                   hmax, rng[0], rng[1], rng[2], thr[0], thr[1], C.byref(out))
        assert rc == 0
        return out

    def match_grid(self, grid, angles, ranges, init_pose, rng, step, thr=(0.0, 0.0), rel_pose=None):
        a, r, p, q = self._scan(angles, ranges, init_pose, rel_pose)
        out = OrcResult()
        rc = self.lib.orc_match_grid(grid.h, _dptr(a), _dptr(r), len(a), _dptr(p), _dptr(q),
                                     rng[0], rng[1], rng[2], step[0], step[1], step[2],
                                     thr[0], thr[1], C.byref(out))
        assert rc == 0
        return out

    def refine(self, grid, angles, ranges, init_pose, rel_pose=None, iterations_max=10,
               convergence_threshold=1e-4, lam=1e-4):
        """ScanMatcherLinearSolver::OptimizePose with a fresh solver (n_processed = iterations)."""
        a, r, p, q = self._scan(angles, ranges, init_pose, rel_pose)
        out = OrcResult()
        lam_c = C.c_double(lam)
        rc = self.lib.orc_refine(grid.h, _dptr(a), _dptr(r), len(a), _dptr(p), _dptr(q),
                                 iterations_max, convergence_threshold, C.byref(lam_c), C.byref(out))
        assert rc == 0
        return out

    def hill_climb(self, grid, angles, ranges, init_pose, rel_pose=None, linear_step=0.1, angular_step=0.1,
                   max_iterations=100, max_num_of_refinements=5, greedy=None):
        """ScanMatcherHillClimbing::OptimizePose over CostSquareError (n_processed = iterations,
        n_ignored = step halvings); None when this checker lacks it."""
        if not hasattr(self.lib, "orc_hill_climb"):
            return None
        dp = C.POINTER(C.c_double)
        self.lib.orc_hill_climb.argtypes = [C.c_void_p, dp, dp, C.c_int, dp, dp, C.c_double, C.c_double,
                                            C.c_int, C.c_int, dp, C.POINTER(OrcResult)]
        a, r, p, q = self._scan(angles, ranges, init_pose, rel_pose)
        out = OrcResult()
        gr = None
        if greedy is not None:
            gr_arr = np.ascontiguousarray(greedy, dtype=np.float64)
            gr = _dptr(gr_arr)
        rc = self.lib.orc_hill_climb(grid.h, _dptr(a), _dptr(r), len(a), _dptr(p), _dptr(q), linear_step,
                                     angular_step, max_iterations, max_num_of_refinements, gr, C.byref(out))
        return out if rc == 0 else None

    def loop_search(self, scan_ids, scan_poses, map_ids, map_scan_min, map_scan_max, map_finished,
                    accum_travel_dist, last_finished_scan_id, last_finished_map_id,
                    travel_dist_threshold, node_dist_threshold, num_of_candidate_nodes):
        """LoopSearcherNearest::Search of the reference. Returns [(query scan node, reference scan
        node, reference local map)] in the reference's order, or None when this checker lacks it."""
        if not hasattr(self.lib, "orc_loop_search"):
            return None
        i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        self.lib.orc_loop_search.argtypes = [C.c_int, i32p, dp, C.c_int, i32p, i32p, i32p, i32p, C.c_double,
                                             C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, i32p, C.c_int]
        si = np.ascontiguousarray(scan_ids, dtype=np.int32)
        sp = np.ascontiguousarray(scan_poses, dtype=np.float64).reshape(-1)
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (map_ids, map_scan_min, map_scan_max, map_finished)]
        cap = max(int(num_of_candidate_nodes), 1)
        out = np.zeros(3 * cap, dtype=np.int32)
        n = self.lib.orc_loop_search(len(si), si.ctypes.data_as(i32p), sp.ctypes.data_as(dp), len(arrs[0]),
                                     *[a.ctypes.data_as(i32p) for a in arrs], accum_travel_dist,
                                     last_finished_scan_id, last_finished_map_id, travel_dist_threshold,
                                     node_dist_threshold, num_of_candidate_nodes, out.ctypes.data_as(i32p), cap)
        if n < 0:
            return None
        return [tuple(int(v) for v in out[3 * i:3 * i + 3]) for i in range(n)]

    def carmen_load(self, text):
        """The reference's CarmenLogReader::Load on a log text; returns a handle for hostapi.carmen_records
        (prefix 'orc_carmen_'), or None when this checker lacks it."""
        if not hasattr(self.lib, "orc_carmen_load"):
            return None
        self.lib.orc_carmen_load.restype = C.c_void_p
        self.lib.orc_carmen_load.argtypes = [C.c_char_p]
        self.lib.orc_carmen_destroy.argtypes = [C.c_void_p]
        return self.lib.orc_carmen_load(text.encode())

    def metric_values_string(self, kind, values):
        """'Values' of a ValueSequence<int> (kind 0), <float> (1), <uint64_t> (2) that observed `values`"""
        if not hasattr(self.lib, "orc_metric_values_string"):
            return None
        self.lib.orc_metric_values_string.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_char_p, C.c_int]
        v = np.ascontiguousarray(values, dtype=np.float64)
        buf = C.create_string_buffer(64 * max(len(v), 1) + 16)
        self.lib.orc_metric_values_string(kind, _dptr(v), len(v), buf, len(buf))
        return buf.value.decode()

    def loop_detector(self, hmax, rng, thr, n_threads=1):
        return OracleLoopDetector(self, hmax, rng, thr, n_threads)


class OracleGrid:
    def __init__(self, oracle, h, shape):
        self.oracle, self.h, self.shape = oracle, h, shape

    def precompute(self, win):
        out = np.empty(self.shape, dtype=np.uint16)
        self.oracle.lib.orc_precompute(self.h, win, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def pyramid(self, hmax):
        out = np.empty((hmax + 1,) + tuple(self.shape), dtype=np.uint16)
        self.oracle.lib.orc_precompute_pyramid(self.h, hmax, out.ctypes.data_as(C.POINTER(C.c_uint16)))
        return out

    def close(self):
        if self.h:
            self.oracle.lib.orc_grid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OracleLoopDetector:
    def __init__(self, oracle, hmax, rng, thr, n_threads):
        self.oracle = oracle
        self.n_threads = n_threads
        self.h = oracle.lib.orc_loopdet_create(hmax, rng[0], rng[1], rng[2], thr[0], thr[1], n_threads)

    def clear_cache(self):
        self.oracle.lib.orc_loopdet_clear_cache(self.h)

    def use_linear_solver(self, iterations_max=10, convergence_threshold=1e-4, initial_lambda=1e-4):
        """Refine detected loops with the reference's default final matcher instead of the
        pass-through one (the best_* window fields of the results are then not meaningful)."""
        self.oracle.lib.orc_loopdet_use_linear_solver(self.h, iterations_max, convergence_threshold,
                                                      initial_lambda)

    def detect(self, grids, map_ids, map_poses, scan_idx, scan_poses, angles, ranges):
        """grids: list[OracleGrid] per query; angles/ranges: (n_scans, n_beams)."""
        nq = len(grids)
        gh = (C.c_void_p * nq)(*[g.h for g in grids])
        mid = np.ascontiguousarray(map_ids, dtype=np.int32)
        sid = np.ascontiguousarray(scan_idx, dtype=np.int32)
        mp = np.ascontiguousarray(map_poses, dtype=np.float64).reshape(nq, 3)
        sp = np.ascontiguousarray(scan_poses, dtype=np.float64).reshape(nq, 3)
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        assert a.ndim == 2 and a.shape == r.shape
        out = (OrcResult * nq)()
        el = C.c_double(0.0)
        i32p = C.POINTER(C.c_int32)
        rc = self.oracle.lib.orc_loopdet_detect(
            self.h, nq, gh, mid.ctypes.data_as(i32p), _dptr(mp), sid.ctypes.data_as(i32p),
            _dptr(sp), a.shape[0], a.shape[1], _dptr(a), _dptr(r), out, C.byref(el))
        assert rc == 0
        return list(out), el.value

    def close(self):
        if self.h:
            self.oracle.lib.orc_loopdet_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CACHE = {}


def load(kind="reference"):
    if kind not in _CACHE:
        _CACHE[kind] = Oracle(kind)
    return _CACHE[kind]


class OracleMapBuilder:
    """The reference's GridMapBuilder (one instance per process, re-initialised here): append(pose, scan)
    rebuilds the latest map like the front end does per scan."""

    def __init__(self, oracle, resolution, patch_size, scans_for_latest_map, usable_range_min, usable_range_max,
                 prob_hit, prob_miss):
        self.lib = oracle.lib
        self.p = self.lib.orc_mapbuilder_create(resolution, patch_size, scans_for_latest_map, usable_range_min,
                                                usable_range_max, prob_hit, prob_miss)

    def append(self, pose, angles, ranges, rel_pose=(0.0, 0.0, 0.0), min_range=0.01, max_range=50.0):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        p = np.ascontiguousarray(pose, dtype=np.float64)
        q = np.ascontiguousarray(rel_pose, dtype=np.float64)
        assert self.lib.orc_mapbuilder_append(self.p, _dptr(p), _dptr(a), _dptr(r), len(a), _dptr(q),
                                              min_range, max_range) == 0

    def latest(self, cap_cells=1 << 22):
        geo = np.zeros(6)
        pose = np.zeros(3)
        dense = np.zeros(cap_cells, dtype=np.uint16)
        alloc = np.zeros(cap_cells // 64, dtype=np.uint8)
        assert self.lib.orc_mapbuilder_latest(self.p, _dptr(geo), _dptr(pose), dense.ctypes.data, cap_cells,
                                              alloc.ctypes.data, len(alloc)) == 0
        rows, cols, bs = int(geo[0]), int(geo[1]), int(geo[2])
        return (dense[:rows * cols].reshape(rows, cols).copy(),
                alloc[:(rows // bs) * (cols // bs)].reshape(rows // bs, cols // bs).copy(), (geo[3], geo[4]), pose, bs)


SLAM_COUNTERS = ("scans_in", "scans_processed", "backend_steps", "backend_steps_with_candidates", "loop_queries",
                 "loops_detected", "optimizations", "degenerations", "optimizer_calls",
                 "t_latest_map", "t_match", "t_append", "t_backend", "t_detect")


class OracleSlam:
    """Same surface as hostapi.SlamPipeline, on the compiled reference."""

    def __init__(self, oracle, settings):
        self.lib = oracle.lib
        v = np.ascontiguousarray(settings, dtype=np.float64)
        self.p = self.lib.orc_slam_create(_dptr(v), len(v))
        assert self.p

    def run(self, angles, ranges, odom_poses, time_stamps, min_range=0.01, max_range=50.0, finish=False):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        o = np.ascontiguousarray(odom_poses, dtype=np.float64)
        t = np.ascontiguousarray(time_stamps, dtype=np.float64)
        assert r.ndim == 2 and r.shape[1] == len(a) and o.shape == (r.shape[0], 3) and len(t) == r.shape[0]
        return self.lib.orc_slam_run(self.p, r.shape[0], r.shape[1], _dptr(a), _dptr(r), _dptr(o), _dptr(t),
                                     min_range, max_range, int(finish))

    def counters(self):
        out = np.zeros(len(SLAM_COUNTERS))
        self.lib.orc_slam_counters(self.p, _dptr(out))
        return dict(zip(SLAM_COUNTERS, out.tolist()))

    def _table(self, what, width):
        n = getattr(self.lib, "orc_slam_num_" + what)(self.p)
        out = np.zeros((n, width))
        if n:
            getattr(self.lib, "orc_slam_" + what)(self.p, _dptr(out))
        return out

    def scan_nodes(self):
        return self._table("scan_nodes", 7)

    def local_maps(self):
        return self._table("local_maps", 10)

    def edges(self):
        return self._table("edges", 7)

    def loops(self):
        return self._table("loops", 6)

    def local_map_cells(self, map_id, block_size=16):
        info = self.local_maps()[map_id]
        rows, cols = int(info[6]), int(info[7])
        dense = np.zeros((rows, cols), dtype=np.uint16)
        alloc = np.zeros((rows // block_size, cols // block_size), dtype=np.uint8)
        rc = self.lib.orc_slam_local_map_cells(self.p, map_id, dense.ctypes.data, dense.size, alloc.ctypes.data,
                                               alloc.size)
        assert rc == 0, rc
        return dense, alloc

    def close(self):
        if self.p:
            self.lib.orc_slam_destroy(self.p)
            self.p = None
