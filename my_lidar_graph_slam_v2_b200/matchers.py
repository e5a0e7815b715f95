"""Host-side mirror of the reference's scan-matcher / loop-detector plugin
interface for the accelerated path, on top of the C ABI (capi.py).

Class and argument names follow the reference:
  ScanMatcherCorrelative(name, lowResolution, rangeX, rangeY, rangeTheta)
      scan_matcher_correlative.hpp:59-84
  ScanMatcherBranchBound(name, nodeHeightMax, rangeX, rangeY, rangeTheta)
      scan_matcher_branch_bound.hpp:108-136
  ScanMatcherGridSearch(name, rangeX, rangeY, rangeTheta, stepX, stepY, stepTheta)
      scan_matcher_grid_search.hpp:45-72
  LoopDetectorBranchBound(name, scanMatcher, scoreThreshold, knownRateThreshold)
      loop_detector_branch_bound.hpp, loop_detector_branch_bound.cpp:59-156

All host arithmetic that decides window sizes, steps and candidate poses uses
the reference's own double expressions (Python floats are IEEE doubles and
math.sin/cos/acos are glibc's, no FMA contraction), so the values handed to
the device are bit-identical to what the reference would compute.
The device does everything between "inputs staged" and "best window index";
this module only prepares inputs and converts indices back to poses.
"""
import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import capi


# --- data types (reference: pose.hpp, sensor/sensor_data.hpp, grid_map.hpp) --
def compound(start, diff):
    """pose.hpp:154-166"""
    s, c = math.sin(start[2]), math.cos(start[2])
    return (c * diff[0] - s * diff[1] + start[0],
            s * diff[0] + c * diff[1] + start[1],
            start[2] + diff[2])


def inverse_compound(start, end):
    """pose.hpp:183-198"""
    s, c = math.sin(start[2]), math.cos(start[2])
    dx, dy = end[0] - start[0], end[1] - start[1]
    return (c * dx + s * dy, -s * dx + c * dy, end[2] - start[2])


def move_backward(end, diff):
    """pose.hpp:211-224"""
    theta = end[2] - diff[2]
    s, c = math.sin(theta), math.cos(theta)
    return (end[0] - c * diff[0] + s * diff[1], end[1] - s * diff[0] - c * diff[1], theta)


@dataclass
class ScanData:
    """Sensor::ScanData<double> (sensor/sensor_data.hpp:69-89): beam angles,
    ranges and the sensor pose relative to the robot."""
    angles: np.ndarray
    ranges: np.ndarray
    relative_sensor_pose: tuple = (0.0, 0.0, 0.0)

    def num_of_scans(self):
        return len(self.ranges)


@dataclass
class GridMap:
    """Dense view of Mapping::GridMap (grid_map_new/grid_map.hpp:27): row-major
    u16 values (0 = unknown), resolution and position offset of cell (0, 0).
    The adapter flattens the block-sparse reference map cell by cell
    (GridMap::CopyValues is unusable for 16-bit buffers, grid_map.cpp:343-349)."""
    values: np.ndarray
    resolution: float
    pos_offset: tuple
    map_id: int = -1


@dataclass
class ScanMatchingSummary:
    """scan_matcher.hpp:56-83 (cost / covariance are filled by the CPU epilogue
    of the C++ adapter; here they stay None)."""
    pose_found: bool
    map_local_initial_pose: tuple
    estimated_pose: tuple
    best_sensor_pose: tuple
    result: capi.CsmResult
    win: tuple = (0, 0, 0)
    step: tuple = (0.0, 0.0, 0.0)
    normalized_cost: float = None
    estimated_covariance: np.ndarray = None


def compute_search_step(resolution, ranges):
    """scan_matcher_correlative.cpp:255-274 / scan_matcher_branch_bound.cpp:293-312"""
    max_range = float(np.max(ranges))
    theta = resolution / max_range
    return resolution, resolution, math.acos(1.0 - 0.5 * theta * theta)


def search_window(rng, step):
    """scan_matcher_correlative.cpp:141-146"""
    return (int(math.ceil(0.5 * rng[0] / step[0])),
            int(math.ceil(0.5 * rng[1] / step[1])),
            int(math.ceil(0.5 * rng[2] / step[2])))


def grid_search_offsets(radius, step):
    """The accumulating loop `for (d = -r; d <= r; d += s)` of
    scan_matcher_grid_search.cpp:118-120."""
    out = []
    d = -radius
    while d <= radius:
        out.append(d)
        d += step
    return np.asarray(out, dtype=np.float64)


class _Base:
    _next_map_id = 1 << 40

    def __init__(self, name, handle=None, device=0):
        self.name = name
        self.handle = handle if handle is not None else capi.Handle(device)

    def Name(self):
        return self.name

    def _ensure_map(self, grid_map):
        """Upload (or re-use) the map on the device. Maps with map_id >= 0 are
        cached by id like the FPGA matcher does (scan_matcher_correlative_fpga.cpp:
        261-270); anonymous maps are uploaded on every call (front-end latest map)."""
        if grid_map.map_id >= 0:
            key = grid_map.map_id
            if key in getattr(self, "_uploaded", {}):
                return key, False
        else:
            key = _Base._next_map_id
        self.handle.upload_grid(key, grid_map.values, grid_map.resolution,
                                grid_map.pos_offset[0], grid_map.pos_offset[1])
        if grid_map.map_id >= 0:
            self.__dict__.setdefault("_uploaded", {})[key] = True
        return key, True


class ScanMatcherCorrelative(_Base):
    """Real-time correlative matcher on the GPU (scan_matcher_correlative.cpp:92-244)."""

    def __init__(self, name, low_resolution, range_x, range_y, range_theta, handle=None, device=0):
        super().__init__(name, handle, device)
        self.low_resolution = int(low_resolution)
        self.range = (range_x, range_y, range_theta)

    def compute_coarser_map(self, map_key):
        """ComputeCoarserMap -> PrecomputeGridMap (scan_matcher_correlative.cpp:247-252)"""
        self.handle.build_coarse(map_key, self.low_resolution)

    def optimize_pose(self, grid_map, scan, map_local_initial_pose, score_thr=0.0, known_thr=0.0):
        key, fresh = self._ensure_map(grid_map)
        self.compute_coarser_map(key)
        sensor = compound(map_local_initial_pose, scan.relative_sensor_pose)
        step = compute_search_step(grid_map.resolution, scan.ranges)
        win = search_window(self.range, step)
        r = self.handle.match_rt(key, scan.angles, scan.ranges, sensor, self.low_resolution,
                                 win, step, (score_thr, known_thr))
        best = (sensor[0] + r.best_x * step[0], sensor[1] + r.best_y * step[1],
                sensor[2] + r.best_t * step[2])
        est = move_backward(best, scan.relative_sensor_pose)
        return ScanMatchingSummary(bool(r.found), tuple(map_local_initial_pose), est, best, r, win, step)


class ScanMatcherBranchBound(_Base):
    """Branch-and-bound matcher on the GPU (scan_matcher_branch_bound.cpp:87-278)."""

    def __init__(self, name, node_height_max, range_x, range_y, range_theta, handle=None, device=0):
        super().__init__(name, handle, device)
        self.node_height_max = int(node_height_max)
        self.range = (range_x, range_y, range_theta)

    def compute_coarser_maps(self, map_key):
        """ComputeCoarserMaps -> PrecomputeGridMaps (scan_matcher_branch_bound.cpp:281-290)"""
        self.handle.build_pyramid(map_key, self.node_height_max)

    def optimize_pose(self, grid_map, scan, map_local_initial_pose, score_thr=0.0, known_thr=0.0):
        key, fresh = self._ensure_map(grid_map)
        self.compute_coarser_maps(key)
        sensor = compound(map_local_initial_pose, scan.relative_sensor_pose)
        step = compute_search_step(grid_map.resolution, scan.ranges)
        win = search_window(self.range, step)
        r = self.handle.match_bb(key, scan.angles, scan.ranges, sensor, self.node_height_max,
                                 win, step, (score_thr, known_thr))
        best = (sensor[0] + step[0] * r.best_x, sensor[1] + step[1] * r.best_y,
                sensor[2] + step[2] * r.best_t)
        est = move_backward(best, scan.relative_sensor_pose)
        return ScanMatchingSummary(bool(r.found), tuple(map_local_initial_pose), est, best, r, win, step)


class ScanMatcherGridSearch(_Base):
    """Exhaustive grid search on the GPU (scan_matcher_grid_search.cpp:69-178)."""

    def __init__(self, name, range_x, range_y, range_theta, step_x, step_y, step_theta,
                 handle=None, device=0):
        super().__init__(name, handle, device)
        self.range = (range_x, range_y, range_theta)
        self.step = (step_x, step_y, step_theta)

    def optimize_pose(self, grid_map, scan, map_local_initial_pose, score_thr=0.0, known_thr=0.0):
        key, fresh = self._ensure_map(grid_map)
        sensor = compound(map_local_initial_pose, scan.relative_sensor_pose)
        dx = grid_search_offsets(self.range[0] / 2.0, self.step[0])
        dy = grid_search_offsets(self.range[1] / 2.0, self.step[1])
        dt = grid_search_offsets(self.range[2] / 2.0, self.step[2])
        r = self.handle.match_grid(key, scan.angles, scan.ranges, sensor, dx, dy, dt,
                                   (score_thr, known_thr))
        if r.found:
            best = (sensor[0] + dx[r.best_x], sensor[1] + dy[r.best_y], sensor[2] + dt[r.best_t])
        else:
            best = sensor
        est = move_backward(best, scan.relative_sensor_pose)
        return ScanMatchingSummary(bool(r.found), tuple(map_local_initial_pose), est, best, r,
                                   (0, 0, 0), self.step)


@dataclass
class LoopDetectionQuery:
    """loop_detector.hpp:27-55 with references resolved to plain data."""
    scan: ScanData
    scan_id: int                  # id of the scan data (uploaded once per batch)
    scan_global_pose: tuple       # mQueryScanNode.mGlobalPose
    local_map: GridMap            # mReferenceLocalMap.mMap, map_id = LocalMapId
    local_map_global_pose: tuple  # mReferenceLocalMapNode.mGlobalPose
    scan_node_id: int = 0


@dataclass
class LoopDetectionResult:
    """loop_detector.hpp:65-92 (coarse stage: relative pose before the CPU refiner)."""
    relative_pose: tuple
    local_map_pose: tuple
    local_map_id: int
    scan_node_id: int
    result: capi.CsmResult
    query_index: int


class LoopDetectorBranchBound(_Base):
    """Batched, GPU-resident LoopDetectorBranchBound::Detect
    (loop_detector_branch_bound.cpp:59-156): per query, the pyramid of the
    local map is built on first touch and cached by LocalMapId, the initial
    pose is InverseCompound(map pose, scan pose), and a result is emitted only
    when the coarse match clears both thresholds."""

    def __init__(self, name, scan_matcher, score_threshold, known_rate_threshold):
        assert 0.0 < score_threshold <= 1.0          # loop_detector_branch_bound.cpp:54-55
        assert 0.0 < known_rate_threshold <= 1.0
        super().__init__(name, scan_matcher.handle)
        self.scan_matcher = scan_matcher
        self.score_threshold = score_threshold
        self.known_rate_threshold = known_rate_threshold
        self._cached_maps = set()
        self._cached_scans = {}
        self.last_best_key = None

    def clear_cache(self):
        self._cached_maps.clear()

    def prepare(self, queries):
        """Upload what the batch needs and return the csm_loop_query array."""
        h = self.handle
        hmax = self.scan_matcher.node_height_max
        new_maps = []
        for q in queries:
            mid = q.local_map.map_id
            assert mid >= 0, "loop detection maps are identified by LocalMapId"
            if mid not in self._cached_maps:
                h.upload_grid(mid, q.local_map.values, q.local_map.resolution,
                              q.local_map.pos_offset[0], q.local_map.pos_offset[1])
                self._cached_maps.add(mid)
                new_maps.append(mid)
            if self._cached_scans.get(q.scan_id) is not q.scan:
                h.upload_scan(q.scan_id, q.scan.angles, q.scan.ranges)
                self._cached_scans[q.scan_id] = q.scan
        if new_maps:
            h.build_pyramids(new_maps, hmax)
        arr = (capi.CsmLoopQuery * len(queries))()
        steps = {}
        for i, q in enumerate(queries):
            init = inverse_compound(q.local_map_global_pose, q.scan_global_pose)
            sensor = compound(init, q.scan.relative_sensor_pose)
            skey = (q.scan_id, q.local_map.resolution)
            if skey not in steps:
                st = compute_search_step(q.local_map.resolution, q.scan.ranges)
                steps[skey] = (st, search_window(self.scan_matcher.range, st))
            step, win = steps[skey]
            a = arr[i]
            a.map_id, a.scan_id = q.local_map.map_id, q.scan_id
            a.sensor_pose[0], a.sensor_pose[1], a.sensor_pose[2] = sensor
            a.win_x, a.win_y, a.win_t = win
            a.step_x, a.step_y, a.step_t = step
            a.score_thr, a.known_thr = self.score_threshold, self.known_rate_threshold
        return arr

    def detect(self, queries, query_index_base=0):
        arr = self.prepare(queries)
        hmax = self.scan_matcher.node_height_max
        res = self.handle.loop_batch(arr, len(queries), hmax, query_index_base)
        out = []
        for i, (q, r) in enumerate(zip(queries, res)):
            if not r.found:
                continue
            a = arr[i]
            best = (a.sensor_pose[0] + a.step_x * r.best_x, a.sensor_pose[1] + a.step_y * r.best_y,
                    a.sensor_pose[2] + a.step_t * r.best_t)
            est = move_backward(best, q.scan.relative_sensor_pose)
            out.append(LoopDetectionResult(est, tuple(q.local_map_global_pose), q.local_map.map_id,
                                           q.scan_node_id, r, i))
        return out, res
