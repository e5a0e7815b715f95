"""Seeded synthetic occupancy submaps and LiDAR scans.

Grids are rectangular rooms with 3-cell-thick walls, a few rectangular pillars
and partial inner walls: wall cells ~ U[45000, 65534], free interior cells
~ U[1, 16000], everything outside the room unknown (0). Values are capped at
65534 because the reference lookup table is one entry short at 65535
(grid_values.cpp:32-35). Scans are ray-cast analytically
against the wall centre lines from a true pose inside the room, with Gaussian
range noise and one beam forced to `rmax` so that the angular search step
acos(1 - 0.5 (res / rmax)^2) is fixed (scan_matcher_correlative.cpp:255-274).

Everything here is numpy on the host; it feeds the CPU checkers and the CUDA
path with identical bytes.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Room:
    """Axis-aligned room in map-local metres: outer box + obstacle boxes."""
    x0: float
    y0: float
    x1: float
    y1: float
    boxes: list = field(default_factory=list)      # (bx0, by0, bx1, by1) pillars
    stubs: list = field(default_factory=list)      # thin inner walls (segments)

    def segments(self):
        segs = [(self.x0, self.y0, self.x1, self.y0), (self.x1, self.y0, self.x1, self.y1),
                (self.x1, self.y1, self.x0, self.y1), (self.x0, self.y1, self.x0, self.y0)]
        for bx0, by0, bx1, by1 in self.boxes:
            segs += [(bx0, by0, bx1, by0), (bx1, by0, bx1, by1),
                     (bx1, by1, bx0, by1), (bx0, by1, bx0, by0)]
        segs += list(self.stubs)
        return np.asarray(segs, dtype=np.float64)


@dataclass
class Submap:
    grid: np.ndarray          # (rows, cols) uint16, 0 = unknown
    res: float
    off_x: float
    off_y: float
    room: Room


def make_room(rng, width=16.0, height=12.0, jitter=2.0, n_boxes=4, n_stubs=3):
    w = width + rng.uniform(-jitter, jitter)
    h = height + rng.uniform(-jitter, jitter)
    room = Room(-w / 2, -h / 2, w / 2, h / 2)
    for _ in range(n_boxes):
        bw, bh = rng.uniform(0.4, 1.6, size=2)
        cx = rng.uniform(room.x0 + 1.5, room.x1 - 1.5 - bw)
        cy = rng.uniform(room.y0 + 1.5, room.y1 - 1.5 - bh)
        room.boxes.append((cx, cy, cx + bw, cy + bh))
    for _ in range(n_stubs):
        if rng.random() < 0.5:      # stub attached to the bottom/top wall
            x = rng.uniform(room.x0 + 2.0, room.x1 - 2.0)
            ln = rng.uniform(1.0, 0.35 * h)
            if rng.random() < 0.5:
                room.stubs.append((x, room.y0, x, room.y0 + ln))
            else:
                room.stubs.append((x, room.y1, x, room.y1 - ln))
        else:
            y = rng.uniform(room.y0 + 2.0, room.y1 - 2.0)
            ln = rng.uniform(1.0, 0.35 * w)
            if rng.random() < 0.5:
                room.stubs.append((room.x0, y, room.x0 + ln, y))
            else:
                room.stubs.append((room.x1, y, room.x1 - ln, y))
    return room


def rasterize(room, rng, rows=512, cols=512, res=0.05, wall_cells=3, off_jitter=True):
    """Room -> dense u16 grid centred in a rows x cols map."""
    off_x = -0.5 * cols * res
    off_y = -0.5 * rows * res
    if off_jitter:              # non-round offset so floor() sees generic fractions
        off_x += rng.uniform(-0.5, 0.5) * res
        off_y += rng.uniform(-0.5, 0.5) * res
    grid = np.zeros((rows, cols), dtype=np.uint16)

    def cidx(x):
        return int(np.floor((x - off_x) / res))

    def ridx(y):
        return int(np.floor((y - off_y) / res))

    # free interior
    r0, r1 = ridx(room.y0), ridx(room.y1)
    c0, c1 = cidx(room.x0), cidx(room.x1)
    grid[r0:r1 + 1, c0:c1 + 1] = rng.integers(1, 16001, size=(r1 - r0 + 1, c1 - c0 + 1),
                                              dtype=np.uint16)
    # pillars are solid: unknown inside, wall on the faces
    for bx0, by0, bx1, by1 in room.boxes:
        grid[ridx(by0):ridx(by1) + 1, cidx(bx0):cidx(bx1) + 1] = 0
    half = wall_cells // 2
    for x0, y0, x1, y1 in room.segments():
        ra, rb = sorted((ridx(y0), ridx(y1)))
        ca, cb = sorted((cidx(x0), cidx(x1)))
        ra, rb = max(ra - half, 0), min(rb + half, rows - 1)
        ca, cb = max(ca - half, 0), min(cb + half, cols - 1)
        grid[ra:rb + 1, ca:cb + 1] = rng.integers(
            45000, 65535, size=(rb - ra + 1, cb - ca + 1), dtype=np.uint16)
    assert grid.max() <= 65534
    return Submap(grid, res, off_x, off_y, room)


def raycast(room, pose, n_beams=360, sigma=0.01, rmax=11.40, rng=None, force_rmax=True):
    """Analytic ray casting against the room's wall centre lines."""
    angles = -np.pi + 2.0 * np.pi * np.arange(n_beams) / n_beams
    th = pose[2] + angles
    dx, dy = np.cos(th), np.sin(th)
    segs = room.segments()
    ax, ay = segs[:, 0][None, :], segs[:, 1][None, :]
    ex, ey = segs[:, 2][None, :] - ax, segs[:, 3][None, :] - ay
    # ray: p + t d ; segment: a + u e  ->  solve with cross products
    den = dx[:, None] * ey - dy[:, None] * ex
    wx, wy = ax - pose[0], ay - pose[1]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (wx * ey - wy * ex) / den
        u = (wx * dy[:, None] - wy * dx[:, None]) / den
    ok = (np.abs(den) > 1e-12) & (t > 1e-6) & (u >= 0.0) & (u <= 1.0)
    t = np.where(ok, t, np.inf)
    ranges = t.min(axis=1)
    ranges = np.where(np.isfinite(ranges), ranges, rmax)
    if rng is not None and sigma > 0:
        ranges = ranges + rng.normal(0.0, sigma, size=n_beams)
    ranges = np.clip(ranges, 0.05, rmax)
    if force_rmax:
        ranges[n_beams // 7] = rmax
    return angles.astype(np.float64), ranges.astype(np.float64)


def random_pose_in_room(room, rng, margin=2.5):
    for _ in range(1000):
        x = rng.uniform(room.x0 + margin, room.x1 - margin)
        y = rng.uniform(room.y0 + margin, room.y1 - margin)
        inside_box = any(bx0 - 0.5 <= x <= bx1 + 0.5 and by0 - 0.5 <= y <= by1 + 0.5
                         for bx0, by0, bx1, by1 in room.boxes)
        if not inside_box:
            return np.array([x, y, rng.uniform(-np.pi, np.pi)])
    raise RuntimeError("no free pose found")


@dataclass
class MatchCase:
    submap: Submap
    angles: np.ndarray
    ranges: np.ndarray
    true_pose: np.ndarray
    init_pose: np.ndarray


def make_match_case(seed, rows=512, cols=512, res=0.05, n_beams=360, rmax=11.40,
                    offset=(0.15, 0.15, 0.05), room_size=(16.0, 12.0)):
    """One (submap, scan, initial pose) triple. `offset` bounds the uniform
    perturbation of the initial pose (60 % of the half window in the configs)."""
    rng = np.random.default_rng(seed)
    room = make_room(rng, room_size[0], room_size[1])
    sub = rasterize(room, rng, rows, cols, res)
    true_pose = random_pose_in_room(room, rng)
    angles, ranges = raycast(room, true_pose, n_beams, 0.01, rmax, rng)
    init = true_pose + rng.uniform(-1.0, 1.0, size=3) * np.asarray(offset)
    return MatchCase(sub, angles, ranges, true_pose, init)


# --- BASELINE.json configurations -----------------
DEG = np.pi / 180.0
CFG1 = dict(name="cfg1_rt", low_res=5, rng=(0.5, 0.5, 10.0 * DEG), thr=(0.0, 0.0),
            rows=512, cols=512, res=0.05, n_beams=360, offset=(0.15, 0.15, 3.0 * DEG))
CFG2 = dict(name="cfg2_bb", hmax=5, rng=(2.0, 2.0, 30.0 * DEG), thr=(0.0, 0.0),
            rows=512, cols=512, res=0.05, n_beams=360, offset=(0.6, 0.6, 9.0 * DEG))
CFG3 = dict(name="cfg3_loop", hmax=6, rng=(2.5, 2.5, 0.5), thr=(0.55, 0.6),
            rows=512, cols=512, res=0.05, n_beams=360, offset=(0.75, 0.75, 0.15),
            n_maps=256, true_fraction=0.25)
CFG4 = dict(name="cfg4_grid", rng=(4.0, 4.0, 60.0 * DEG), step=(0.025, 0.025, 0.1 * DEG),
            thr=(0.0, 0.0), rows=1280, cols=1280, res=0.025, n_beams=1080,
            offset=(1.2, 1.2, 18.0 * DEG), rmax=11.40)


def case_for(cfg, seed):
    return make_match_case(seed, cfg["rows"], cfg["cols"], cfg["res"], cfg["n_beams"],
                           cfg.get("rmax", 11.40), cfg["offset"], cfg.get("room_size", (16.0, 12.0)))


@dataclass
class LoopBatch:
    submaps: list             # n_maps Submap
    map_ids: np.ndarray       # (nq,) int32
    map_poses: np.ndarray     # (nq, 3) global pose of each local map node
    scan_poses: np.ndarray    # (nq, 3) global pose of the query scan node
    scan_idx: np.ndarray      # (nq,) index into angles/ranges
    angles: np.ndarray        # (n_scans, n_beams)
    ranges: np.ndarray
    is_true: np.ndarray       # (nq,) bool, map drawn from the scan's room


def make_loop_batch(seed, n_maps=256, true_fraction=0.25, rows=512, cols=512, res=0.05,
                    n_beams=360, offset=(0.75, 0.75, 0.15), map_id_base=0):
    """One query scan against n_maps candidate submaps (BASELINE.json configs[2]).

    About `true_fraction` of the submaps are re-rasterised (fresh cell values)
    from the room the scan was taken in; the rest are other rooms."""
    rng = np.random.default_rng(seed)
    room = make_room(rng)
    true_pose = random_pose_in_room(room, rng)
    angles, ranges = raycast(room, true_pose, n_beams, 0.01, 11.40, rng)
    submaps, is_true = [], np.zeros(n_maps, dtype=bool)
    map_poses = np.zeros((n_maps, 3))
    scan_poses = np.zeros((n_maps, 3))
    for m in range(n_maps):
        is_true[m] = rng.random() < true_fraction
        r = room if is_true[m] else make_room(rng)
        submaps.append(rasterize(r, rng, rows, cols, res))
        # global pose of the map node, and of the scan node such that
        # InverseCompound(map, scan) = true pose + bounded perturbation
        mp = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-np.pi, np.pi)])
        local = true_pose + rng.uniform(-1.0, 1.0, size=3) * np.asarray(offset)
        c, s = np.cos(mp[2]), np.sin(mp[2])
        sp = np.array([c * local[0] - s * local[1] + mp[0],
                       s * local[0] + c * local[1] + mp[1], mp[2] + local[2]])
        map_poses[m], scan_poses[m] = mp, sp
    return LoopBatch(submaps, np.arange(n_maps, dtype=np.int32) + map_id_base, map_poses,
                     scan_poses, np.zeros(n_maps, dtype=np.int32), angles[None, :].copy(),
                     ranges[None, :].copy(), is_true)


def dense_to_blocks(grid, log2bs=4):
    """Block-sparse form of a dense grid, as the reference stores it
    (grid_map.cpp:262-266, 522-535: a block is allocated once any of its cells
    has been written): the blocks that hold a non-zero cell, row-major inside a
    block, and their positions block_row * block_cols + block_col."""
    bs = 1 << log2bs
    rows, cols = grid.shape
    assert rows % bs == 0 and cols % bs == 0
    br, bc = rows // bs, cols // bs
    tiles = grid.reshape(br, bs, bc, bs).swapaxes(1, 2)            # (br, bc, bs, bs)
    used = tiles.reshape(br, bc, -1).any(axis=2)
    index = np.flatnonzero(used.reshape(-1)).astype(np.int32)
    blocks = np.ascontiguousarray(tiles.reshape(br * bc, bs, bs)[index])
    return blocks, index, br, bc


def make_pose_graph_summary(seed, n_maps=12, scans_per_map=(6, 14), loop=True):
    """A pose-graph summary for the loop searcher: scan nodes along a closed
    trajectory with odometry-like jitter, consecutive runs of them grouped into local maps (all
    finished), the accumulated travel distance of the walk.
    Returns dict(scan_ids, scan_poses, map_ids, map_scan_min, map_scan_max, map_finished,
    accum_travel_dist, last_finished_scan_id, last_finished_map_id)."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(scans_per_map[0], scans_per_map[1] + 1, size=n_maps)
    n = int(counts.sum())
    t = np.linspace(0.0, (2.0 if loop else 1.2) * np.pi, n)
    radius = rng.uniform(6.0, 14.0)
    xs = radius * np.cos(t) * rng.uniform(0.7, 1.3) + rng.normal(0.0, 0.05, n)
    ys = radius * np.sin(t) + rng.normal(0.0, 0.05, n)
    th = t + np.pi / 2 + rng.normal(0.0, 0.02, n)
    poses = np.stack([xs, ys, th], axis=1)
    first_id = int(rng.integers(0, 5))
    scan_ids = np.arange(first_id, first_id + n, dtype=np.int32)
    ends = np.cumsum(counts)
    starts = ends - counts
    accum = float(np.sum(np.hypot(np.diff(xs), np.diff(ys))))
    return dict(scan_ids=scan_ids, scan_poses=poses, map_ids=np.arange(n_maps, dtype=np.int32),
                map_scan_min=scan_ids[starts], map_scan_max=scan_ids[ends - 1],
                map_finished=np.ones(n_maps, dtype=np.int32), accum_travel_dist=accum,
                last_finished_scan_id=int(scan_ids[-1]), last_finished_map_id=n_maps - 1)


# ---- BASELINE configs[4]: a closed-loop trajectory in a corridor world ----------------

def corridor_world(rng, width=60.0, height=40.0, corridor=5.0, stub_every=4.0):
    """A 60 x 40 m block with a corridor running around its inner core: outer wall, core as one big
    pillar, and short wall stubs every few metres on alternating sides so that a 11.4 m scanner always
    sees something that fixes its position along the corridor."""
    world = Room(0.0, 0.0, width, height)
    world.boxes.append((corridor, corridor, width - corridor, height - corridor))
    side = 0
    for x in np.arange(corridor + 2.0, width - corridor - 2.0, stub_every):
        ln = rng.uniform(0.6, 1.6)
        xs = x + rng.uniform(-0.8, 0.8)
        for y_wall, sgn, y_core in ((0.0, 1.0, corridor), (height, -1.0, height - corridor)):
            if side % 2 == 0:
                world.stubs.append((xs, y_wall, xs, y_wall + sgn * ln))
            else:
                world.stubs.append((xs, y_core, xs, y_core - sgn * ln))
            side += 1
    for y in np.arange(corridor + 2.0, height - corridor - 2.0, stub_every):
        ln = rng.uniform(0.6, 1.6)
        ys = y + rng.uniform(-0.8, 0.8)
        for x_wall, sgn, x_core in ((0.0, 1.0, corridor), (width, -1.0, width - corridor)):
            if side % 2 == 0:
                world.stubs.append((x_wall, ys, x_wall + sgn * ln, ys))
            else:
                world.stubs.append((x_core, ys, x_core - sgn * ln, ys))
            side += 1
    return world


def corridor_path(world, corridor=5.0, radius=1.5):
    """Centre line of the corridor as a closed polyline with rounded corners: (points (n, 2), length)."""
    m = 0.5 * corridor
    x0, y0, x1, y1 = world.x0 + m, world.y0 + m, world.x1 - m, world.y1 - m
    pts = []
    corners = [(x1, y0, -0.5 * np.pi), (x1, y1, 0.0), (x0, y1, 0.5 * np.pi), (x0, y0, np.pi)]
    start = [(x0 + radius, y0)]
    pts += start
    for cx, cy, a0 in corners:
        # arc centre sits `radius` inside the corner
        ccx = cx - radius if cx == x1 else cx + radius
        ccy = cy - radius if cy == y1 else cy + radius
        for a in np.linspace(a0, a0 + 0.5 * np.pi, 13):
            pts.append((ccx + radius * np.cos(a), ccy + radius * np.sin(a)))
    pts.append(start[0])
    pts = np.asarray(pts)
    seg = np.hypot(*np.diff(pts, axis=0).T)
    return pts, float(seg.sum())


def corridor_trajectory(world, n_scans, spacing, rng, n_beams=360, sigma=0.01, rmax=11.40, corridor=5.0,
                        odom_sigma=(0.002, 0.002, 0.0005), odom_bias=(0.0005, 0.0, 0.0001)):
    """n_scans true poses `spacing` metres apart along the corridor loop (several laps when the path is
    shorter than the trip), their scans, and an odometry track that integrates the true steps with a
    bias and noise (Carmen-style drifting odometry). Returns dict(true, odom, angles, ranges, stamps)."""
    pts, length = corridor_path(world, corridor)
    seg = np.hypot(*np.diff(pts, axis=0).T)
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    s = (np.arange(n_scans) * spacing) % length
    idx = np.clip(np.searchsorted(cum, s, side="right") - 1, 0, len(seg) - 1)
    f = (s - cum[idx]) / seg[idx]
    xy = pts[idx] + f[:, None] * (pts[idx + 1] - pts[idx])
    head = np.arctan2(pts[idx + 1, 1] - pts[idx, 1], pts[idx + 1, 0] - pts[idx, 0])
    true = np.column_stack([xy, head])
    ranges = np.empty((n_scans, n_beams))
    angles = None
    for k in range(n_scans):
        angles, ranges[k] = raycast(world, true[k], n_beams, sigma, rmax, rng, force_rmax=(k == 0))
    # every scan must carry the same longest range, or the angular search step would change from scan to scan
    ranges[:, n_beams // 7] = rmax
    odom = np.zeros((n_scans, 3))
    bias, sig = np.asarray(odom_bias), np.asarray(odom_sigma)
    for k in range(1, n_scans):
        c, sn = np.cos(true[k - 1, 2]), np.sin(true[k - 1, 2])
        d = true[k, :2] - true[k - 1, :2]
        rel = np.array([c * d[0] + sn * d[1], -sn * d[0] + c * d[1],
                        (true[k, 2] - true[k - 1, 2] + np.pi) % (2 * np.pi) - np.pi])
        rel = rel + bias + rng.normal(0.0, 1.0, 3) * sig
        c, sn = np.cos(odom[k - 1, 2]), np.sin(odom[k - 1, 2])
        odom[k] = [odom[k - 1, 0] + c * rel[0] - sn * rel[1], odom[k - 1, 1] + sn * rel[0] + c * rel[1],
                   odom[k - 1, 2] + rel[2]]
    return dict(true=true, odom=odom, angles=angles, ranges=ranges, stamps=0.1 * np.arange(n_scans))
