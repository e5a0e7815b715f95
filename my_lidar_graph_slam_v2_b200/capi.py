"""ctypes binding of the C ABI in include/csm_b200.h (libcsm_b200.so).

This is the same boundary a C++ adapter deriving from the reference's
ScanMatcher / LoopDetector classes binds to. There is no
fallback: if the CUDA library is missing or no device is usable, loading or
csm_create raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

CSM_OK = 0
FLAG_FP_MARGIN, FLAG_KEY_TIE, FLAG_EDGE, FLAG_EXACT = 1, 2, 4, 8


class CsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("csm error %d: %s" % (code, msg))
        self.code = code


class CsmResult(C.Structure):
    _fields_ = [
        ("found", C.c_int32), ("best_x", C.c_int32), ("best_y", C.c_int32), ("best_t", C.c_int32),
        ("sum_value", C.c_int64), ("n_known", C.c_int32), ("flags", C.c_int32),
        ("normalized_score", C.c_double), ("n_processed", C.c_int32), ("n_ignored", C.c_int32),
    ]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class CsmRefineParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("reserved", C.c_int32), ("convergence_threshold", C.c_double),
                ("lambda_", C.c_double), ("covariance_scale", C.c_double)]


class CsmRefined(C.Structure):
    _fields_ = [("pose", C.c_double * 3), ("covariance", C.c_double * 9), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("lambda_", C.c_double), ("iterations", C.c_int32),
                ("valid", C.c_int32)]


class CsmRefineQuery(C.Structure):
    _fields_ = [("map_id", C.c_int64), ("scan_id", C.c_int64), ("sensor_pose", C.c_double * 3)]


class CsmLoopQuery(C.Structure):
    _fields_ = [
        ("map_id", C.c_int64), ("scan_id", C.c_int64), ("sensor_pose", C.c_double * 3),
        ("win_x", C.c_int32), ("win_y", C.c_int32), ("win_t", C.c_int32), ("reserved", C.c_int32),
        ("step_x", C.c_double), ("step_y", C.c_double), ("step_t", C.c_double),
        ("score_thr", C.c_double), ("known_thr", C.c_double),
    ]


EXPORTS = [
    "csm_version", "csm_device_count", "csm_create", "csm_destroy", "csm_last_error",
    "csm_stream", "csm_synchronize", "csm_launch_count", "csm_set_option", "csm_alloc_pinned", "csm_free_pinned",
    "csm_upload_grid", "csm_upload_grid_device", "csm_upload_grids", "csm_upload_grid_blocks",
    "csm_upload_grids_blocks", "csm_release_grid", "csm_build_coarse",
    "csm_build_pyramid", "csm_build_pyramids", "csm_drop_pyramids", "csm_download_level",
    "csm_upload_scan", "csm_release_scan", "csm_match_rt", "csm_match_bb", "csm_match_grid",
    "csm_loop_batch_enqueue", "csm_loop_batch_finish", "csm_loop_batch",
    "csm_set_refiner", "csm_loop_batch_finish_refined", "csm_refine_batch", "csm_set_epilogue", "csm_last_epilogue", "csm_share_copy_stream",
    "csm_best_key_device", "csm_decode_best_key", "csm_debug_frontier_counts", "csm_debug_timings",
    "csm_debug_bound_level", "csm_exact_rerun_count", "csm_debug_node_list",
    "csm_detect_step_enqueue", "csm_comm_unique_id", "csm_comm_init_rank", "csm_comm_init_all",
    "csm_comm_allreduce_best", "csm_comm_allreduce_best_all", "csm_comm_best_result", "csm_comm_destroy",
    "csm_comm_allreduce_word", "csm_comm_allreduce_words_all",
    "csm_map_set_update_tables", "csm_map_create", "csm_map_resize", "csm_map_reset_values",
    "csm_map_insert_rays", "csm_map_download_allocation", "csm_map_download_cells",
]

_LIB = None


def comm_unique_id():
    """128 bytes that identify a new communicator (rank 0 creates them, every rank gets them)."""
    buf = C.create_string_buffer(128)
    rc = load().csm_comm_unique_id(buf)
    if rc != CSM_OK:
        raise CsmError(rc, "csm_comm_unique_id (libnccl.so.2 not found?)")
    return buf.raw


def comm_init_all(handles):
    """One process, several GPUs: a communicator over `handles` (one per device)."""
    arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
    rc = load().csm_comm_init_all(arr, len(handles))
    if rc != CSM_OK:
        raise CsmError(rc, handles[0].last_error())


def comm_allreduce_best_all(handles):
    arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
    tickets = (C.c_int * len(handles))()
    rc = load().csm_comm_allreduce_best_all(arr, len(handles), tickets)
    if rc != CSM_OK:
        raise CsmError(rc, handles[0].last_error())
    return list(tickets)


def lib_path():
    return os.environ.get("CSM_B200_LIB", _build.LIB)


def load():
    """Load libcsm_b200.so (built in-tree by build.py). Raises if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise FileNotFoundError(
            "%s not built: run `python -m my_lidar_graph_slam_v2_b200.build` "
            "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    dp, u16p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_uint16), C.POINTER(C.c_int64)
    H = C.c_void_p
    rp = C.POINTER(CsmResult)
    lib.csm_version.restype = C.c_int
    lib.csm_device_count.restype = C.c_int
    lib.csm_create.argtypes = [C.c_int, C.c_uint, C.POINTER(H)]
    lib.csm_destroy.argtypes = [H]
    lib.csm_last_error.argtypes = [H]
    lib.csm_last_error.restype = C.c_char_p
    lib.csm_stream.argtypes = [H]
    lib.csm_stream.restype = C.c_void_p
    lib.csm_synchronize.argtypes = [H]
    lib.csm_launch_count.argtypes = [H]
    lib.csm_launch_count.restype = C.c_int64
    lib.csm_set_option.argtypes = [H, C.c_char_p, C.c_int]
    lib.csm_alloc_pinned.argtypes = [C.c_size_t]
    lib.csm_alloc_pinned.restype = C.c_void_p
    lib.csm_free_pinned.argtypes = [C.c_void_p]
    lib.csm_upload_grid.argtypes = [H, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    lib.csm_upload_grid_device.argtypes = lib.csm_upload_grid.argtypes
    lib.csm_upload_grids.argtypes = [H, C.c_int, i64p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_double, dp, dp]
    i32p = C.POINTER(C.c_int32)
    lib.csm_upload_grid_blocks.argtypes = [H, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_double, C.c_double, C.c_double]
    lib.csm_upload_grids_blocks.argtypes = [H, C.c_int, i64p, C.c_void_p, C.c_void_p, i32p, C.c_int, C.c_int,
                                            C.c_int, C.c_double, dp, dp]
    lib.csm_release_grid.argtypes = [H, C.c_int64]
    lib.csm_build_coarse.argtypes = [H, C.c_int64, C.c_int]
    lib.csm_build_pyramid.argtypes = [H, C.c_int64, C.c_int]
    lib.csm_build_pyramids.argtypes = [H, C.c_int, i64p, C.c_int]
    lib.csm_drop_pyramids.argtypes = [H, C.c_int, i64p]
    lib.csm_download_level.argtypes = [H, C.c_int64, C.c_int, u16p]
    lib.csm_upload_scan.argtypes = [H, C.c_int64, dp, dp, C.c_int]
    lib.csm_release_scan.argtypes = [H, C.c_int64]
    scan = [H, C.c_int64, dp, dp, C.c_int, dp]
    lib.csm_match_rt.argtypes = scan + [C.c_int] * 4 + [C.c_double] * 5 + [rp]
    lib.csm_match_bb.argtypes = scan + [C.c_int] * 4 + [C.c_double] * 5 + [rp]
    lib.csm_match_grid.argtypes = scan + [dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_double, C.c_double, rp]
    lq = C.POINTER(CsmLoopQuery)
    lib.csm_loop_batch_enqueue.argtypes = [H, lq, C.c_int, C.c_int, C.c_int]
    lib.csm_loop_batch_finish.argtypes = [H, rp, C.c_int]
    lib.csm_loop_batch.argtypes = [H, lq, C.c_int, C.c_int, C.c_int, rp]
    if hasattr(lib, "csm_detect_step_enqueue"): lib.csm_detect_step_enqueue.argtypes = [H, i64p, C.c_int, C.c_int, lq, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    if hasattr(lib, "csm_comm_unique_id"): lib.csm_comm_unique_id.argtypes = [C.c_void_p]
    if hasattr(lib, "csm_comm_init_rank"): lib.csm_comm_init_rank.argtypes = [H, C.c_void_p, C.c_int, C.c_int]
    if hasattr(lib, "csm_comm_init_all"): lib.csm_comm_init_all.argtypes = [C.POINTER(H), C.c_int]
    if hasattr(lib, "csm_comm_allreduce_best"): lib.csm_comm_allreduce_best.argtypes = [H, C.POINTER(C.c_int)]
    if hasattr(lib, "csm_comm_allreduce_best_all"): lib.csm_comm_allreduce_best_all.argtypes = [C.POINTER(H), C.c_int, C.POINTER(C.c_int)]
    if hasattr(lib, "csm_comm_best_result"): lib.csm_comm_best_result.argtypes = [H, C.c_int, C.POINTER(C.c_uint64)]
    if hasattr(lib, "csm_comm_allreduce_word"): lib.csm_comm_allreduce_word.argtypes = [H, C.c_uint64, C.POINTER(C.c_int)]
    if hasattr(lib, "csm_comm_allreduce_words_all"): lib.csm_comm_allreduce_words_all.argtypes = [C.POINTER(H), C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    if hasattr(lib, "csm_comm_destroy"): lib.csm_comm_destroy.argtypes = [H]
    lib.csm_set_refiner.argtypes = [H, C.POINTER(CsmRefineParams)]
    lib.csm_loop_batch_finish_refined.argtypes = [H, rp, C.POINTER(CsmRefined), C.c_int]
    lib.csm_refine_batch.argtypes = [H, C.POINTER(CsmRefineQuery), C.c_int, C.POINTER(CsmRefineParams),
                                     C.POINTER(CsmRefined)]
    lib.csm_share_copy_stream.argtypes = [H, H]
    lib.csm_set_epilogue.argtypes = [H, C.c_double]
    lib.csm_last_epilogue.argtypes = [H, C.POINTER(CsmRefined)]
    lib.csm_debug_frontier_counts.argtypes = [H, C.POINTER(C.c_uint)]
    if hasattr(lib, "csm_debug_node_list"): lib.csm_debug_node_list.argtypes = [H, C.c_int, C.POINTER(C.c_uint64), C.c_int]
    if hasattr(lib, "csm_exact_rerun_count"): lib.csm_exact_rerun_count.restype = C.c_int64
    if hasattr(lib, "csm_exact_rerun_count"): lib.csm_exact_rerun_count.argtypes = [H]
    if hasattr(lib, "csm_debug_bound_level"): lib.csm_debug_bound_level.argtypes = [H, C.c_int64, C.c_int, C.POINTER(C.c_uint8)]
    lib.csm_debug_timings.argtypes = [H, C.c_char_p, C.c_size_t, C.POINTER(C.c_float), C.c_int]
    lib.csm_best_key_device.argtypes = [H]
    lib.csm_best_key_device.restype = C.c_void_p
    lib.csm_decode_best_key.argtypes = [C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    _LIB = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Handle:
    """One csm_handle: one CUDA stream + device-resident map / scan caches."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.csm_create(device, 0, C.byref(h))
        if rc != CSM_OK:
            raise CsmError(rc, "csm_create failed (no usable CUDA device %d; there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    @classmethod
    def from_pointer(cls, ptr, device=0):
        """Non-owning view of a csm_handle created elsewhere (e.g. by the C++ host classes)."""
        self = cls.__new__(cls)
        self.lib = load()
        self.h = C.c_void_p(ptr)
        self.device = device
        self.borrowed = True
        return self

    def last_error(self):
        return self.lib.csm_last_error(self.h).decode()

    def _check(self, rc):
        if rc != CSM_OK:
            raise CsmError(rc, self.last_error())

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "borrowed", False):
                self.lib.csm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.lib.csm_stream(self.h)

    def synchronize(self):
        self._check(self.lib.csm_synchronize(self.h))

    def set_option(self, name, value):
        self._check(self.lib.csm_set_option(self.h, name.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.csm_launch_count(self.h))

    # -- maps -------------------------------------------------------------
    def upload_grid(self, map_id, dense, res, off_x, off_y):
        dense = np.ascontiguousarray(dense, dtype=np.uint16)
        self._check(self.lib.csm_upload_grid(self.h, map_id, dense.ctypes.data, dense.shape[0],
                                             dense.shape[1], res, off_x, off_y))

    def upload_grid_ptr(self, map_id, ptr, rows, cols, res, off_x, off_y, device=False):
        fn = self.lib.csm_upload_grid_device if device else self.lib.csm_upload_grid
        self._check(fn(self.h, map_id, ptr, rows, cols, res, off_x, off_y))

    def upload_grids_ptr(self, map_ids, ptrs, rows, cols, res, off_x, off_y):
        """map_ids: int64 array, ptrs: (c_void_p * n) of host buffers, off_x/off_y: float64 arrays."""
        self._check(self.lib.csm_upload_grids(self.h, len(map_ids), map_ids.ctypes.data_as(C.POINTER(C.c_int64)),
                                              ptrs, rows, cols, res, _dptr(off_x), _dptr(off_y)))

    def upload_grid_blocks(self, map_id, blocks, block_index, log2bs, block_rows, block_cols, res, off_x, off_y):
        """blocks: (n, bs, bs) uint16, block_index: (n,) int32 = block_row * block_cols + block_col."""
        blocks = np.ascontiguousarray(blocks, dtype=np.uint16)
        block_index = np.ascontiguousarray(block_index, dtype=np.int32)
        self._check(self.lib.csm_upload_grid_blocks(self.h, map_id, blocks.ctypes.data, block_index.ctypes.data,
                                                    len(block_index), log2bs, block_rows, block_cols,
                                                    res, off_x, off_y))

    def upload_grids_blocks_ptr(self, map_ids, blocks_ptr, index_ptr, counts, log2bs, block_rows, block_cols,
                                res, off_x, off_y):
        """Batched block-sparse upload from raw host pointers (pinned for async copies)."""
        self._check(self.lib.csm_upload_grids_blocks(
            self.h, len(map_ids), map_ids.ctypes.data_as(C.POINTER(C.c_int64)), blocks_ptr, index_ptr,
            counts.ctypes.data_as(C.POINTER(C.c_int32)), log2bs, block_rows, block_cols, res,
            _dptr(off_x), _dptr(off_y)))

    def release_grid(self, map_id):
        self._check(self.lib.csm_release_grid(self.h, map_id))

    def build_coarse(self, map_id, win):
        self._check(self.lib.csm_build_coarse(self.h, map_id, win))

    def build_pyramid(self, map_id, hmax):
        self._check(self.lib.csm_build_pyramid(self.h, map_id, hmax))

    def build_pyramids(self, map_ids, hmax):
        ids = np.ascontiguousarray(map_ids, dtype=np.int64)
        self._check(self.lib.csm_build_pyramids(self.h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int64)), hmax))

    def drop_pyramids(self, map_ids):
        ids = np.ascontiguousarray(map_ids, dtype=np.int64)
        self._check(self.lib.csm_drop_pyramids(self.h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int64))))

    def download_level(self, map_id, level, shape):
        out = np.empty(shape, dtype=np.uint16)
        self._check(self.lib.csm_download_level(self.h, map_id, level, out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    # -- scans ------------------------------------------------------------
    def upload_scan(self, scan_id, angles, ranges):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        self._check(self.lib.csm_upload_scan(self.h, scan_id, _dptr(a), _dptr(r), len(a)))

    # -- matchers ---------------------------------------------------------
    def match_rt(self, map_id, angles, ranges, sensor_pose, low_res, win, step, thr):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        p = np.ascontiguousarray(sensor_pose, dtype=np.float64)
        out = CsmResult()
        self._check(self.lib.csm_match_rt(self.h, map_id, _dptr(a), _dptr(r), len(a), _dptr(p), low_res,
                                          win[0], win[1], win[2], step[0], step[1], step[2],
                                          thr[0], thr[1], C.byref(out)))
        return out

    def match_bb(self, map_id, angles, ranges, sensor_pose, hmax, win, step, thr):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        p = np.ascontiguousarray(sensor_pose, dtype=np.float64)
        out = CsmResult()
        self._check(self.lib.csm_match_bb(self.h, map_id, _dptr(a), _dptr(r), len(a), _dptr(p), hmax,
                                          win[0], win[1], win[2], step[0], step[1], step[2],
                                          thr[0], thr[1], C.byref(out)))
        return out

    def match_grid(self, map_id, angles, ranges, sensor_pose, dx, dy, dt, thr):
        a = np.ascontiguousarray(angles, dtype=np.float64)
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        p = np.ascontiguousarray(sensor_pose, dtype=np.float64)
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        dy = np.ascontiguousarray(dy, dtype=np.float64)
        dt = np.ascontiguousarray(dt, dtype=np.float64)
        out = CsmResult()
        self._check(self.lib.csm_match_grid(self.h, map_id, _dptr(a), _dptr(r), len(a), _dptr(p),
                                            _dptr(dx), len(dx), _dptr(dy), len(dy), _dptr(dt), len(dt),
                                            thr[0], thr[1], C.byref(out)))
        return out

    # -- loop detection -----------------------------------------------------
    def loop_batch_enqueue(self, queries, nq, hmax, query_index_base=0):
        self._check(self.lib.csm_loop_batch_enqueue(self.h, queries, nq, hmax, query_index_base))

    def loop_batch_finish(self, nq, results=None):
        results = results if results is not None else (CsmResult * nq)()
        self._check(self.lib.csm_loop_batch_finish(self.h, results, nq))
        return results

    def loop_batch(self, queries, nq, hmax, query_index_base=0):
        self.loop_batch_enqueue(queries, nq, hmax, query_index_base)
        return self.loop_batch_finish(nq)

    def detect_step_enqueue(self, map_ids, queries, nq, hmax, query_index_base=0, drop=True):
        """The whole first-touch step in one call (csm_detect_step_enqueue): rebuild what the search reads
        above level 0 for `map_ids`, enqueue the batch, start the best-word exchange when the handle has a
        communicator. Returns the exchange ticket (-1: none)."""
        ids = np.ascontiguousarray(map_ids, dtype=np.int64)
        ticket = C.c_int(-1)
        self._check(self.lib.csm_detect_step_enqueue(self.h, ids.ctypes.data_as(C.POINTER(C.c_int64)), len(ids),
                                                     1 if drop else 0, queries, nq, hmax, query_index_base,
                                                     C.byref(ticket)))
        return ticket.value

    # -- exchange of the packed best word over NCCL (the library calls NCCL itself) ------------------
    def comm_init_rank(self, id128, rank, world):
        buf = C.create_string_buffer(bytes(id128), 128)
        self._check(self.lib.csm_comm_init_rank(self.h, buf, rank, world))

    def comm_allreduce_best(self):
        ticket = C.c_int(-1)
        self._check(self.lib.csm_comm_allreduce_best(self.h, C.byref(ticket)))
        return ticket.value

    def comm_allreduce_word(self, word):
        ticket = C.c_int(-1)
        self._check(self.lib.csm_comm_allreduce_word(self.h, C.c_uint64(word), C.byref(ticket)))
        return ticket.value

    def comm_best_result(self, ticket):
        word = C.c_uint64(0)
        self._check(self.lib.csm_comm_best_result(self.h, ticket, C.byref(word)))
        return int(word.value)

    # -- refinement (ScanMatcherLinearSolver on the device) ------------------------
    def set_refiner(self, max_iterations=10, convergence_threshold=1e-4, lambda_=1e-4, covariance_scale=1e4,
                    enabled=True):
        """Loop batches enqueued from now on refine the poses they find (csm_set_refiner)."""
        if not enabled:
            self._check(self.lib.csm_set_refiner(self.h, None))
            return
        p = CsmRefineParams(max_iterations, 0, convergence_threshold, lambda_, covariance_scale)
        self._check(self.lib.csm_set_refiner(self.h, C.byref(p)))

    def loop_batch_finish_refined(self, nq, results=None, refined=None):
        results = results if results is not None else (CsmResult * nq)()
        refined = refined if refined is not None else (CsmRefined * nq)()
        self._check(self.lib.csm_loop_batch_finish_refined(self.h, results, refined, nq))
        return results, refined

    def refine_batch(self, jobs, max_iterations=10, convergence_threshold=1e-4, lambda_=1e-4,
                     covariance_scale=1e4):
        """jobs: [(map_id, scan_id, (x, y, theta))] -> CsmRefined array."""
        n = len(jobs)
        q = (CsmRefineQuery * n)()
        for i, (mid, sid, pose) in enumerate(jobs):
            q[i].map_id, q[i].scan_id = int(mid), int(sid)
            q[i].sensor_pose[:] = [float(v) for v in pose]
        p = CsmRefineParams(max_iterations, 0, convergence_threshold, lambda_, covariance_scale)
        out = (CsmRefined * n)()
        self._check(self.lib.csm_refine_batch(self.h, q, n, C.byref(p), out))
        return out

    def set_epilogue(self, covariance_scale):
        """> 0: csm_match_rt / csm_match_bb also compute cost and covariance at the decided pose."""
        self._check(self.lib.csm_set_epilogue(self.h, float(covariance_scale)))

    def last_epilogue(self):
        out = CsmRefined()
        self._check(self.lib.csm_last_epilogue(self.h, C.byref(out)))
        return out

    def bound_level(self, map_id, level, shape):
        """Level `level` of the map's u8 bound levels (csm_bounds.cuh), untiled to (rows, cols)."""
        out = np.empty(shape, dtype=np.uint8)
        self._check(self.lib.csm_debug_bound_level(self.h, map_id, level, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def node_list(self, level, cap=1 << 24):
        """(q, t, xi, yi) of the nodes in list(level) after the last batch (see option "bb_stop_level")."""
        out = np.empty(cap, dtype=np.uint64)
        n = self.lib.csm_debug_node_list(self.h, level, out.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
        if n < 0:
            self._check(n)
        w = out[:n]
        return np.stack([(w >> np.uint64(48)) & np.uint64(0xffff), (w >> np.uint64(32)) & np.uint64(0xffff),
                         (w >> np.uint64(16)) & np.uint64(0xffff), w & np.uint64(0xffff)], axis=1).astype(np.int64)

    def exact_rerun_count(self):
        return int(self.lib.csm_exact_rerun_count(self.h))

    def frontier_counts(self):
        out = (C.c_uint * 8)()
        self._check(self.lib.csm_debug_frontier_counts(self.h, out))
        return list(out)

    def timings(self):
        """[(phase name, ms)] of the last loop batch / pyramid build (option "timing" must be 1)."""
        names = C.create_string_buffer(8192)
        ms = (C.c_float * 256)()
        n = self.lib.csm_debug_timings(self.h, names, 8192, ms, 256)
        return list(zip(names.value.decode().split(";")[:n], [ms[i] for i in range(n)]))

    def best_key_device_ptr(self):
        return self.lib.csm_best_key_device(self.h)

    def decode_best_key(self, word):
        k, q = C.c_int64(), C.c_int32()
        self.lib.csm_decode_best_key(C.c_uint64(int(word)), C.byref(k), C.byref(q))
        return k.value, q.value
