"""Parity of the CUDA path (through the C ABI) against the CPU checkers and the
committed golden vectors. Bar: best pose index, integer score and found flag
bit-exact; the double score within 1e-5 relative (it is in fact bit-identical:
the device re-sums in the reference's order)."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import as_matchers_read, assert_match, grid_of, load_golden, sha
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth

pytestmark = pytest.mark.gpu

CFGS = {"CFG1": synth.CFG1, "CFG2": synth.CFG2, "CFG3": synth.CFG3}


def _scan(case):
    return matchers.ScanData(case.angles, case.ranges)


# --------------------------------------------------------------------------
# precomputation
# --------------------------------------------------------------------------
@pytest.mark.parametrize("entry", load_golden("reference_vectors.json")["pyramids"],
                         ids=lambda e: "%dx%d" % (e["rows"], e["cols"]))
def test_pyramid_golden(handle, entry):
    rows, cols = entry["rows"], entry["cols"]
    rng = np.random.default_rng(entry["seed"])
    if rows >= 256:
        grid = synth.rasterize(synth.make_room(rng, 8.0, 6.0, 1.0), rng, rows, cols, 0.05).grid
    else:
        grid = rng.integers(0, 65535, size=(rows, cols), dtype=np.uint16)
        grid[rng.random((rows, cols)) < 0.5] = 0
    assert sha(grid) == entry["grid_sha"], "synthetic generator drifted"
    handle.upload_grid(77, grid, 0.05, -1.0, -2.0)
    handle.build_pyramid(77, 6)
    for h in range(7):
        assert sha(handle.download_level(77, h, (rows, cols))) == entry["levels"][h], "level %d" % h
    for w, digest in entry["coarse"].items():
        handle.build_coarse(77, int(w))
        assert sha(handle.download_level(77, -int(w), (rows, cols))) == digest, "win %s" % w
    handle.release_grid(77)


@pytest.mark.parametrize("shape", [(512, 512), (128, 320), (48, 16), (16, 16)])
def test_pyramid_vs_checker(handle, checker, shape):
    rng = np.random.default_rng(hash(shape) % 1000)
    grid = rng.integers(0, 65535, size=shape, dtype=np.uint16)
    grid[rng.random(shape) < 0.6] = 0
    g = checker.grid(grid, 0.05, 0.0, 0.0)
    handle.upload_grid(78, grid, 0.05, 0.0, 0.0)
    for hmax in (3, 6):
        handle.drop_pyramids([78])
        handle.build_pyramid(78, hmax)
        ref = g.pyramid(hmax)
        for h in range(hmax + 1):
            assert np.array_equal(handle.download_level(78, h, shape), ref[h]), (hmax, h)
    handle.release_grid(78)


@pytest.mark.parametrize("shape", [(512, 512), (48, 16), (16, 208), (144, 272), (32, 32)])
def test_coarse_map_windows_vs_checker(handle, checker, shape):
    """PrecomputeGridMap(map, win) for windows of every kind: the tiled separable kernel (2..32),
    the generic one (1, > 32), windows larger than the map."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    grid = rng.integers(0, 65535, size=shape, dtype=np.uint16)
    grid[rng.random(shape) < 0.7] = 0
    g = checker.grid(grid, 0.05, 0.0, 0.0)
    handle.upload_grid(79, grid, 0.05, 0.0, 0.0)
    for win in (1, 2, 3, 5, 8, 17, 32, 33, 40):
        handle.build_coarse(79, win)
        got = handle.download_level(79, -win, shape)
        assert np.array_equal(got, g.precompute(win)), "win %d: %d cells differ" % (win, int((got != g.precompute(win)).sum()))
    handle.release_grid(79)


@pytest.mark.parametrize("shape,hmax", [((64, 64), 6), ((512, 512), 6), ((32, 128), 4), ((128, 16), 6),
                                        ((16, 512), 3), ((192, 336), 5), ((512, 512), 1), ((48, 48), 2),
                                        ((256, 512), 6), ((384, 128), 6), ((144, 64), 6), ((272, 256), 5)])
@pytest.mark.parametrize("mode", [2, 3])
def test_pyramid_streaming_kernel(handle, checker, shape, hmax, mode):
    """The single-pass streaming builders (used for big batches) against the checker: mode 2 takes
    the register-ring kernel when rows % 32 == 0 (else the shared-memory-ring one), mode 3 always
    the shared-memory-ring kernel."""
    handle.set_option("pyramid_mode", mode)
    try:
        ids, grids = [], []
        for k in range(3):
            rng = np.random.default_rng(1200 + k + shape[0])
            grid = rng.integers(0, 65535, size=shape, dtype=np.uint16)
            grid[rng.random(shape) < (0.2 + 0.3 * k)] = 0
            handle.upload_grid(300 + k, grid, 0.05, 0.0, 0.0)
            ids.append(300 + k)
            grids.append(grid)
        handle.build_pyramids(ids, hmax)
        for i, grid in zip(ids, grids):
            ref = checker.grid(grid, 0.05, 0.0, 0.0).pyramid(hmax)
            for h in range(hmax + 1):
                got = handle.download_level(i, h, shape)
                assert np.array_equal(got, ref[h]), "map %d level %d: %d cells differ" % (
                    i, h, int((got != ref[h]).sum()))
            handle.release_grid(i)
    finally:
        handle.set_option("pyramid_mode", 0)


def test_pyramid_batch(handle, checker):
    ids, grids = [], []
    for k in range(5):
        rng = np.random.default_rng(900 + k)
        shape = [(64, 64), (96, 32), (64, 64), (32, 128), (64, 64)][k]
        grid = rng.integers(0, 65535, size=shape, dtype=np.uint16)
        grid[rng.random(shape) < 0.7] = 0
        handle.upload_grid(100 + k, grid, 0.05, 0.0, 0.0)
        ids.append(100 + k)
        grids.append(grid)
    handle.build_pyramids(ids, 4)
    for i, grid in zip(ids, grids):
        ref = checker.grid(grid, 0.05, 0.0, 0.0).pyramid(4)
        for h in range(5):
            assert np.array_equal(handle.download_level(i, h, grid.shape), ref[h])
        handle.release_grid(i)


# --------------------------------------------------------------------------
# single-scan matchers against the golden vectors of the reference
# --------------------------------------------------------------------------
def _run_golden(handle, m):
    case = synth.case_for(synth.CFG1, m["seed"])
    assert sha(case.submap.grid) == m["grid_sha"] and sha(case.ranges) == m["scan_sha"], \
        "synthetic generator drifted"
    gm, scan = grid_of(case), _scan(case)
    thr = tuple(m["thr"])
    if m["kind"] == "rt":
        rng = tuple(m.get("rng", synth.CFG1["rng"]))
        mt = matchers.ScanMatcherCorrelative("rt", m["low_res"], *rng, handle=handle)
    elif m["kind"] == "bb":
        mt = matchers.ScanMatcherBranchBound("bb", m["hmax"], *m["rng"], handle=handle)
    else:
        mt = matchers.ScanMatcherGridSearch("gs", *m["rng"], *m["step"], handle=handle)
    return mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)


@pytest.mark.parametrize("m", load_golden("reference_vectors.json")["matches"],
                         ids=lambda m: "%s-%d-%s" % (m["kind"], m["seed"], m["thr"][0]))
def test_match_golden(handle, m):
    s = _run_golden(handle, m)
    e = m["expect"]
    assert_match(s.result, e, "%s seed %d" % (m["kind"], m["seed"]))
    assert s.result.flags & capi.FLAG_FP_MARGIN == 0
    if e["found"]:
        for a, b in zip(s.estimated_pose, e["est_pose"]):
            assert a == b, "estimated pose not bit-identical"
    if m["kind"] == "rt":
        # the replay reproduces the reference's own processed / ignored counters
        assert (s.result.n_processed, s.result.n_ignored) == (e["n_processed"], e["n_ignored"])


# --------------------------------------------------------------------------
# live comparison with the checker on fresh seeds
# --------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(2000, 2008))
def test_rt_vs_checker(handle, checker, seed):
    case = synth.case_for(synth.CFG1, seed)
    gm, scan = grid_of(case), _scan(case)
    g = checker.grid(case.submap.grid, case.submap.res, case.submap.off_x, case.submap.off_y)
    for low_res, thr in ((5, (0.0, 0.0)), (5, (0.5, 0.55)), (4, (0.0, 0.0)), (1, (0.2, 0.1))):
        mt = matchers.ScanMatcherCorrelative("rt", low_res, *synth.CFG1["rng"], handle=handle)
        s = mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)
        o = checker.match_rt(g, case.angles, case.ranges, case.init_pose, low_res, synth.CFG1["rng"], thr)
        assert_match(s.result, o, "rt seed %d low_res %d thr %s" % (seed, low_res, thr))
        assert (s.result.n_processed, s.result.n_ignored) == (o.n_processed, o.n_ignored)


@pytest.mark.parametrize("seed", range(2100, 2108))
def test_bb_vs_checker(handle, checker, seed):
    case = synth.case_for(synth.CFG2, seed)
    gm, scan = grid_of(case), _scan(case)
    g = checker.grid(case.submap.grid, case.submap.res, case.submap.off_x, case.submap.off_y)
    for hmax, rng, thr in ((5, synth.CFG2["rng"], (0.0, 0.0)), (6, synth.CFG3["rng"], synth.CFG3["thr"]),
                           (3, (1.0, 0.6, 0.2), (0.3, 0.3)), (0, (0.3, 0.3, 0.05), (0.0, 0.0))):
        mt = matchers.ScanMatcherBranchBound("bb", hmax, *rng, handle=handle)
        s = mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)
        o = checker.match_bb(g, case.angles, case.ranges, case.init_pose, hmax, rng, thr)
        assert_match(s.result, o, "bb seed %d hmax %d" % (seed, hmax))


def test_bb_unreachable_threshold(handle, checker):
    case = synth.case_for(synth.CFG2, 2199)
    gm, scan = grid_of(case), _scan(case)
    mt = matchers.ScanMatcherBranchBound("bb", 5, *synth.CFG2["rng"], handle=handle)
    s = mt.optimize_pose(gm, scan, tuple(case.init_pose), 0.999, 0.0)
    g = checker.grid(case.submap.grid, case.submap.res, case.submap.off_x, case.submap.off_y)
    o = checker.match_bb(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"], (0.999, 0.0))
    assert o.found == 0 and s.result.found == 0
    assert (s.result.best_x, s.result.best_y, s.result.best_t) == (o.best_x, o.best_y, o.best_t) == (0, 0, 0)


@pytest.mark.parametrize("seed", range(2200, 2204))
def test_grid_vs_checker(handle, checker, seed):
    case = synth.case_for(synth.CFG1, seed)
    gm, scan = grid_of(case), _scan(case)
    g = checker.grid(case.submap.grid, case.submap.res, case.submap.off_x, case.submap.off_y)
    for rng, step, thr in (((0.5, 0.5, 0.08), (0.05, 0.05, 0.004), (0.0, 0.0)),      # integer-shift path
                           ((0.4, 0.6, 0.05), (0.1, 0.05, 0.005), (0.4, 0.5)),       # stride-2 offsets
                           ((0.3, 0.3, 0.04), (0.03, 0.07, 0.005), (0.0, 0.0))):     # per-candidate FP path
        mt = matchers.ScanMatcherGridSearch("gs", *rng, *step, handle=handle)
        s = mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)
        o = checker.match_grid(g, case.angles, case.ranges, case.init_pose, rng, step, thr)
        assert_match(s.result, o, "grid seed %d step %s" % (seed, step))


@pytest.mark.parametrize("ndx,ndy", [(33, 9), (34, 21), (36, 5), (66, 11), (97, 7), (129, 3)])
def test_grid_window_remainder_columns(handle, checker, ndx, ndy):
    """Windows of 32 k + r candidate columns (r <= 4): the TMA window kernel scores k chunks per lane
    and the r remainder columns with lanes running over candidate rows. Against the reference and
    against the plain global-memory kernel, with thresholds that leave only part of the window."""
    case = synth.case_for(synth.CFG1, 2600 + ndx)
    gm, scan = grid_of(case), _scan(case)
    g = checker.grid(case.submap.grid, case.submap.res, case.submap.off_x, case.submap.off_y)
    res = case.submap.res
    rng = ((ndx - 1) * res + 1e-9, (ndy - 1) * res + 1e-9, 0.02)
    step = (res, res, 0.005)
    assert len(matchers.grid_search_offsets(rng[0] / 2, step[0])) == ndx
    assert len(matchers.grid_search_offsets(rng[1] / 2, step[1])) == ndy
    mt = matchers.ScanMatcherGridSearch("gs", *rng, *step, handle=handle)
    for thr in ((0.0, 0.0), (0.3, 0.5)):
        handle.set_option("window_mode", 2)
        a = mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)
        handle.set_option("window_mode", 3)
        a16 = mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)
        handle.set_option("window_mode", 1)
        b = mt.optimize_pose(gm, scan, tuple(case.init_pose), *thr)
        handle.set_option("window_mode", 0)
        o = checker.match_grid(g, case.angles, case.ranges, case.init_pose, rng, step, thr)
        assert_match(a.result, o, "tma (32-bit tiles) %dx%d %s" % (ndx, ndy, thr))
        assert_match(a16.result, o, "tma (u16 tiles) %dx%d %s" % (ndx, ndy, thr))
        assert_match(b.result, o, "global %dx%d %s" % (ndx, ndy, thr))


def test_empty_and_out_of_map(handle, checker):
    """All-unknown map: nothing found; scan far outside the map: nothing found."""
    case = synth.case_for(synth.CFG1, 2300)
    empty = matchers.GridMap(np.zeros((64, 64), np.uint16), 0.05, (-1.6, -1.6))
    scan = _scan(case)
    for mt in (matchers.ScanMatcherCorrelative("rt", 5, *synth.CFG1["rng"], handle=handle),
               matchers.ScanMatcherBranchBound("bb", 4, *synth.CFG2["rng"], handle=handle),
               matchers.ScanMatcherGridSearch("gs", 0.3, 0.3, 0.05, 0.05, 0.05, 0.005, handle=handle)):
        assert not mt.optimize_pose(empty, scan, (0.0, 0.0, 0.0)).pose_found
    gm = grid_of(case)
    far = (500.0, -300.0, 0.3)
    g = checker.grid(case.submap.grid, case.submap.res, case.submap.off_x, case.submap.off_y)
    mt = matchers.ScanMatcherCorrelative("rt", 5, *synth.CFG1["rng"], handle=handle)
    s = mt.optimize_pose(gm, scan, far)
    o = checker.match_rt(g, case.angles, case.ranges, far, 5, synth.CFG1["rng"], (0.0, 0.0))
    assert_match(s.result, o, "far away")


def test_single_beam_and_low_edge(handle, checker):
    """N = 1 beam; map whose occupied cells touch row/col 0 (RT path is exact there too)."""
    rng = np.random.default_rng(5)
    grid = rng.integers(1, 65535, size=(64, 64), dtype=np.uint16)
    gm = matchers.GridMap(grid, 0.05, (0.0, 0.0))
    g = checker.grid(grid, 0.05, 0.0, 0.0)
    angles = -np.pi + 2 * np.pi * np.arange(90) / 90
    ranges = np.full(90, 0.9)
    ranges[3] = 1.4
    pose = (0.45, 0.40, 0.2)      # scan reaches negative indices
    mt = matchers.ScanMatcherCorrelative("rt", 4, 0.6, 0.6, 0.3, handle=handle)
    s = mt.optimize_pose(gm, matchers.ScanData(angles, ranges), pose)
    o = checker.match_rt(g, angles, ranges, pose, 4, (0.6, 0.6, 0.3), (0.0, 0.0))
    assert_match(s.result, o, "low edge rt")
    assert (s.result.n_processed, s.result.n_ignored) == (o.n_processed, o.n_ignored)
    one_a, one_r = np.array([0.3]), np.array([1.0])
    s = mt.optimize_pose(gm, matchers.ScanData(one_a, one_r), (1.0, 1.0, 0.0))
    o = checker.match_rt(g, one_a, one_r, (1.0, 1.0, 0.0), 4, (0.6, 0.6, 0.3), (0.0, 0.0))
    assert_match(s.result, o, "single beam", flags_ok=capi.FLAG_KEY_TIE)      # one beam: many candidates tie


def test_bb_edge_flag(handle):
    """Branch-and-bound raises CSM_FLAG_EDGE when a node window straddles row / column 0 (there the
    reference's coarse bound is not admissible) and leaves it clear otherwise."""
    rng = np.random.default_rng(6)
    grid = rng.integers(1, 65535, size=(64, 64), dtype=np.uint16)
    gm = matchers.GridMap(grid, 0.05, (0.0, 0.0))
    angles = -np.pi + 2 * np.pi * np.arange(90) / 90
    ranges = np.full(90, 0.9)
    bb = matchers.ScanMatcherBranchBound("bb", 3, 0.6, 0.6, 0.3, handle=handle)
    r = bb.optimize_pose(gm, matchers.ScanData(angles, ranges), (0.45, 0.40, 0.2)).result
    assert r.flags & capi.FLAG_EDGE
    case = synth.case_for(synth.CFG1, 2100)
    r = matchers.ScanMatcherBranchBound("bb", 5, *synth.CFG2["rng"], handle=handle).optimize_pose(
        grid_of(case), _scan(case), tuple(case.init_pose)).result
    assert r.flags == 0


# --------------------------------------------------------------------------
# bound levels of the sweep, key ties
# --------------------------------------------------------------------------
def _bound_level_numpy(grid, level):
    """ceil(v / 257) followed by the forward 2^level x 2^level maximum, zeros beyond the map."""
    b = ((grid.astype(np.int64) + 256) // 257).astype(np.uint8)
    w = 1 << level
    rows, cols = b.shape
    pad = np.zeros((rows + w, cols + w), dtype=np.uint8)
    pad[:rows, :cols] = b
    out = np.zeros_like(b)
    for dr in range(w):
        for dc in range(w):
            np.maximum(out, pad[dr:dr + rows, dc:dc + cols], out=out)
    return out


@pytest.mark.parametrize("shape", [(512, 512), (128, 320), (48, 16), (100, 38), (16, 16), (264, 130), (37, 70)])
def test_bound_levels_vs_numpy(handle, shape):
    """k_bounds_build: every level equals the u8 encoding of the level-0 cells under the forward
    2^h x 2^h maximum (so 257 * B_h >= the reference's level h, cell for cell), for maps whose extents
    are and are not multiples of the tile / region sizes; the zero padding stays zero."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    grid = rng.integers(0, 65536, size=shape, dtype=np.uint16)
    grid[rng.random(shape) < 0.6] = 0
    grid[0, 0], grid[-1, -1] = 65535, 65535
    handle.upload_grid(81, grid, 0.05, 0.0, 0.0)
    grid = as_matchers_read(grid)          # the library's copy holds 65535 as unknown (k_saturated_unknown)
    assert np.array_equal(handle.download_level(81, 0, shape), grid)
    grid[0, 0], grid[-1, -1] = 65534, 65534
    handle.upload_grid(81, grid, 0.05, 0.0, 0.0)
    for level in (5, 3, 6, 1):
        handle.drop_pyramids([81])
        got = handle.bound_level(81, level, shape)
        for lv in range(level, 0, -1):
            got = handle.bound_level(81, lv, shape)
            exp = _bound_level_numpy(grid, lv)
            assert np.array_equal(got, exp), "level %d of %d: %d cells differ" % (lv, level, int((got != exp).sum()))
    # admissible: 257 * B_h >= the reference's own level h (sliding maximum with the clamped far edge)
    handle.build_pyramid(81, 4)
    ref4 = handle.download_level(81, 4, shape).astype(np.int64)
    b4 = handle.bound_level(81, 4, shape).astype(np.int64)
    inner = (slice(0, max(shape[0] - 16, 0)), slice(0, max(shape[1] - 16, 0)))    # away from the clamped edge
    assert np.all(257 * b4[inner] >= ref4[inner]) and np.all(257 * b4[inner] - ref4[inner] <= 256)
    handle.release_grid(81)


@pytest.mark.parametrize("shape,levels", [((512, 512), 5), ((256, 384), 5), ((512, 512), 6), ((96, 64), 3), ((512, 512), 1)])
def test_bound_levels_from_the_streaming_builder(handle, checker, shape, levels):
    """The streaming pyramid kernel in its bound mode (batches of maps): every level equals the u8
    encoding ceil(v / 257) of the REFERENCE's level (the sliding maximum with the far edge clamped,
    grid_map_builder.cpp:987-1012), cell for cell, and the padding stays zero; the sweep then returns the
    same results from these levels as from the ones k_bounds_build writes."""
    rng = np.random.default_rng(shape[0] + levels)
    ids = list(range(8300, 8309))
    grids = []
    for mid in ids:
        g = rng.integers(0, 65536, size=shape, dtype=np.uint16)
        g[rng.random(shape) < 0.6] = 0
        g[0, 0], g[-1, -1] = 65535, 65535
        handle.upload_grid(mid, g, 0.05, 0.0, 0.0)
        g = as_matchers_read(g)            # what the library keeps: 65535 reads as unknown
        g[0, 0], g[-1, -1] = 65534, 65534
        handle.upload_grid(mid, g, 0.05, 0.0, 0.0)
        grids.append(g)
    handle.set_option("bounds_mode", 2)
    try:
        handle.build_pyramids(ids, levels + 1)
    finally:
        handle.set_option("bounds_mode", 0)
    for mid, g in zip(ids[:3] + ids[-1:], grids[:3] + grids[-1:]):
        ref = checker.grid(g, 0.05, 0.0, 0.0).pyramid(levels)
        for lv in range(1, levels + 1):
            got = handle.bound_level(mid, lv, shape)
            exp = ((ref[lv].astype(np.int64) + 256) // 257).astype(np.uint8)
            assert np.array_equal(got, exp), "map %d level %d: %d cells differ" % (mid, lv, int((got != exp).sum()))
    for mid in ids:
        handle.release_grid(mid)


def test_sweep_over_bound_levels_equals_sweep_over_reference_levels(handle, checker):
    """The same branch-and-bound queries through both sweeps (u8 bound levels / the reference's u16
    levels): identical results field by field (they can differ only in the number of nodes expanded),
    and identical to the reference."""
    fields = ("found", "best_x", "best_y", "best_t", "sum_value", "n_known", "flags", "normalized_score")
    for seed in (2200, 2201, 2202):
        case = synth.case_for(synth.CFG1, seed)
        s = case.submap
        g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
        for hmax, rng, thr in ((6, synth.CFG3["rng"], synth.CFG3["thr"]), (5, synth.CFG2["rng"], (0.3, 0.5)),
                               (1, (0.3, 0.3, 0.05), (0.2, 0.2)), (2, (0.4, 0.4, 0.1), (0.2, 0.2))):
            bb = matchers.ScanMatcherBranchBound("bb", hmax, *rng, handle=handle)
            o = checker.match_bb(g, case.angles, case.ranges, case.init_pose, hmax, rng, thr)
            got = {}
            for bounds in (1, 0):
                handle.set_option("bb_dive", 0)
                handle.set_option("bb_bounds", bounds)
                r = bb.optimize_pose(grid_of(case), _scan(case), tuple(case.init_pose), *thr).result
                assert_match(r, dict(o.asdict(), compare_unfound=False), "seed %d hmax %d bounds %d" % (seed, hmax, bounds))
                got[bounds] = [getattr(r, f) for f in fields]
            handle.set_option("bb_dive", 2)
            handle.set_option("bb_bounds", 1)
            assert got[0] == got[1]


def test_key_tie_flag(handle):
    """Equal integer keys: on a map of one constant value every candidate whose hits stay inside the map
    scores the same. The matchers must say so (CSM_FLAG_KEY_TIE: the reference's pick among equal double
    scores depends on its heap / loop order) and return the documented winner, the first candidate in
    (t, x, y) order for branch-and-bound (whichever sweep runs) and in (y, x, t) order for the grid
    search. A generic map raises no flag."""
    grid = np.full((256, 256), 30000, dtype=np.uint16)
    gm = matchers.GridMap(grid, 0.05, (-6.4, -6.4))
    angles = -np.pi + 2 * np.pi * np.arange(72) / 72
    scan = matchers.ScanData(angles, np.full(72, 2.0))
    bb = matchers.ScanMatcherBranchBound("bb", 3, 0.6, 0.6, 0.2, handle=handle)
    for opts in ((("bb_dive", 0), ("bb_bounds", 1)), (("bb_dive", 0), ("bb_bounds", 0)), (("bb_dive", 1),)):
        for k, v in opts:
            handle.set_option(k, v)
        s = bb.optimize_pose(gm, scan, (0.1, -0.2, 0.3))
        r = s.result
        handle.set_option("bb_dive", 2)
        handle.set_option("bb_bounds", 1)
        assert r.found == 1 and r.flags & capi.FLAG_KEY_TIE, opts
        assert (r.best_x, r.best_y, r.best_t) == (-s.win[0], -s.win[1], -s.win[2]), opts
        assert r.n_known == 72 and r.sum_value == 72 * 30000
    gs = matchers.ScanMatcherGridSearch("gs", 0.4, 0.4, 0.06, 0.05, 0.05, 0.004, handle=handle)
    for mode in (0, 1):
        handle.set_option("window_mode", mode)
        r = gs.optimize_pose(gm, scan, (0.1, -0.2, 0.3)).result
        assert r.found == 1 and r.flags & capi.FLAG_KEY_TIE
        assert (r.best_x, r.best_y, r.best_t) == (0, 0, 0)
    handle.set_option("window_mode", 0)
    # a generic map
    case = synth.case_for(synth.CFG1, 2300)
    r = matchers.ScanMatcherBranchBound("bb", 5, *synth.CFG2["rng"], handle=handle).optimize_pose(
        grid_of(case), _scan(case), tuple(case.init_pose)).result
    assert r.flags == 0


# --------------------------------------------------------------------------
# FP guard band: exact reruns
# --------------------------------------------------------------------------
def test_hit_points_on_cell_boundaries_are_recomputed_exactly(handle, checker):
    """Hit points that sit ON cell boundaries: resolution 2^-4, offsets, sensor position and ranges
    multiples of it, beams along the axes (cos / sin of +-pi/2 are +-6e-17 in libm, so the hit lands a
    hair to either side of the boundary). The projection must flag them and the matchers must still
    return the reference's result, now from host-evaluated indices (CSM_FLAG_EXACT, no
    CSM_FLAG_FP_MARGIN left)."""
    rng = np.random.default_rng(77)
    res = 0.0625
    grid = rng.integers(1, 65535, size=(192, 192), dtype=np.uint16)
    grid[rng.random(grid.shape) < 0.3] = 0
    gm = matchers.GridMap(grid, res, (-6.0, -6.0))
    g = checker.grid(grid, res, -6.0, -6.0)
    angles = np.concatenate([np.array([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi]),
                             -np.pi + 2 * np.pi * np.arange(40) / 40])
    ranges = np.concatenate([np.array([1.0, 1.5, 2.0, 2.5, 0.75]), res * rng.integers(8, 48, size=40)])
    scan = matchers.ScanData(angles, ranges)
    pose = (0.25, -0.5, 0.0)
    n0 = handle.exact_rerun_count()
    rt = matchers.ScanMatcherCorrelative("rt", 4, 0.5, 0.5, 0.2, handle=handle)
    r = rt.optimize_pose(gm, scan, pose).result
    assert r.flags & capi.FLAG_EXACT and not r.flags & capi.FLAG_FP_MARGIN
    assert_match(r, checker.match_rt(g, angles, ranges, pose, 4, (0.5, 0.5, 0.2)), "boundary rt",
                 flags_ok=capi.FLAG_EXACT | capi.FLAG_KEY_TIE)
    for opts in ((("bb_dive", 0), ("bb_bounds", 1)), (("bb_dive", 1),)):
        for k, v in opts:
            handle.set_option(k, v)
        bb = matchers.ScanMatcherBranchBound("bb", 3, 0.5, 0.5, 0.2, handle=handle)
        r = bb.optimize_pose(gm, scan, pose, 0.2, 0.3).result
        handle.set_option("bb_dive", 2)
        assert r.flags & capi.FLAG_EXACT and not r.flags & capi.FLAG_FP_MARGIN
        assert_match(r, checker.match_bb(g, angles, ranges, pose, 3, (0.5, 0.5, 0.2), (0.2, 0.3)), "boundary bb",
                     flags_ok=capi.FLAG_EXACT)
    gs = matchers.ScanMatcherGridSearch("gs", 0.5, 0.5, 0.1, res, res, 0.01, handle=handle)
    r = gs.optimize_pose(gm, scan, pose, 0.2, 0.3).result
    assert r.flags & capi.FLAG_EXACT and not r.flags & capi.FLAG_FP_MARGIN
    assert_match(r, checker.match_grid(g, angles, ranges, pose, (0.5, 0.5, 0.1), (res, res, 0.01), (0.2, 0.3)),
                 "boundary grid", flags_ok=capi.FLAG_EXACT)
    assert handle.exact_rerun_count() >= n0 + 4
    # with the rerun switched off the flag comes through as it is
    handle.set_option("exact_rerun", 0)
    r = rt.optimize_pose(gm, scan, pose).result
    handle.set_option("exact_rerun", 1)
    assert r.flags & capi.FLAG_FP_MARGIN and not r.flags & capi.FLAG_EXACT


def test_poses_off_the_map(handle, checker):
    """The projection takes floor() as a rounded-down add of 1.5 * 2^52 (k_project, project_beam): exact for
    |u| < 2^31 cells, negative coordinates included; beyond 1e9 cells it falls back to the saturating
    conversion. Poses whose beams land left of / below the map (negative cell indices), far away on
    either side, and 2e9 cells away, against the reference. A pose that far has a wide guard band, so the
    exact rerun may answer (CSM_FLAG_EXACT)."""
    case = synth.case_for(synth.CFG1, 2950)
    s = case.submap
    g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
    gm, scan = grid_of(case), _scan(case)
    ok = capi.FLAG_EXACT | capi.FLAG_KEY_TIE
    low = (s.off_x + 1.3, s.off_y + 0.9, 0.4)          # most beams end at negative indices
    poses = [low, (s.off_x - 3.0, s.off_y + 4.0, -1.0), (-512.25, 300.5, 2.0), (731.0, -64.125, 0.1),
             (1.0e8, -1.0e8, 0.3)]
    for pose in poses:
        what = "off-map pose %s" % (pose,)
        r = matchers.ScanMatcherCorrelative("rt", 5, *synth.CFG1["rng"], handle=handle).optimize_pose(
            gm, scan, pose).result
        assert_match(r, checker.match_rt(g, case.angles, case.ranges, pose, 5, synth.CFG1["rng"]), what + " rt",
                     flags_ok=ok)
        gs = matchers.ScanMatcherGridSearch("gs", 0.3, 0.3, 0.04, s.res, s.res, 0.01, handle=handle)
        r = gs.optimize_pose(gm, scan, pose, 0.0, 0.0).result
        assert_match(r, checker.match_grid(g, case.angles, case.ranges, pose, (0.3, 0.3, 0.04),
                                           (s.res, s.res, 0.01), (0.0, 0.0)), what + " grid", flags_ok=ok)
        if pose not in poses[:2]:      # at the low edges the reference's own bound is inadmissible (CSM_FLAG_EDGE)
            bb = matchers.ScanMatcherBranchBound("bb", 4, 1.0, 1.0, 0.2, handle=handle)
            r = bb.optimize_pose(gm, scan, pose, 0.1, 0.1).result
            assert_match(r, checker.match_bb(g, case.angles, case.ranges, pose, 4, (1.0, 1.0, 0.2), (0.1, 0.1)),
                         what + " bb", flags_ok=ok | capi.FLAG_EDGE)


def test_exact_rerun_path_agrees_with_the_reference_everywhere(handle, checker):
    """The exact path is an implementation of its own (host libm + the reference's per-candidate
    arithmetic, exhaustive over the lattice): with the guard band blown up so that every query is
    flagged, all matchers and a loop batch with refinement still return the reference's results."""
    handle.set_option("fp_margin_scale", 10 ** 9)
    try:
        n0 = handle.exact_rerun_count()
        for seed in (2400, 2401):
            case = synth.case_for(synth.CFG1, seed)
            s = case.submap
            g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
            gm, scan = grid_of(case), _scan(case)
            r = matchers.ScanMatcherCorrelative("rt", 5, *synth.CFG1["rng"], handle=handle).optimize_pose(
                gm, scan, tuple(case.init_pose)).result
            assert r.flags == capi.FLAG_EXACT
            assert_match(r, checker.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"]),
                         "rt %d" % seed, flags_ok=capi.FLAG_EXACT)
            r = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=handle).optimize_pose(
                gm, scan, tuple(case.init_pose), *synth.CFG3["thr"]).result
            assert r.flags == capi.FLAG_EXACT
            o = checker.match_bb(g, case.angles, case.ranges, case.init_pose, 6, synth.CFG3["rng"], synth.CFG3["thr"])
            assert_match(r, dict(o.asdict(), compare_unfound=False), "bb %d" % seed, flags_ok=capi.FLAG_EXACT)
            r = matchers.ScanMatcherGridSearch("gs", 0.4, 0.4, 0.06, 0.05, 0.05, 0.004, handle=handle).optimize_pose(
                gm, scan, tuple(case.init_pose)).result
            assert r.flags == capi.FLAG_EXACT
            assert_match(r, checker.match_grid(g, case.angles, case.ranges, case.init_pose, (0.4, 0.4, 0.06),
                                               (0.05, 0.05, 0.004)), "grid %d" % seed, flags_ok=capi.FLAG_EXACT)
        assert handle.exact_rerun_count() == n0 + 6
        # loop batch: coarse results and refined poses
        batch = synth.make_loop_batch(3600, n_maps=12, true_fraction=0.5, map_id_base=9500)
        bb = matchers.ScanMatcherBranchBound("bb", 6, *synth.CFG3["rng"], handle=handle)
        det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
        queries = _loop_queries(batch)
        arr = det.prepare(queries)
        handle.build_pyramids(np.asarray(batch.map_ids, dtype=np.int64), 6)
        handle.set_refiner(10, 1e-4, 1e-4, 1e4)
        handle.loop_batch_enqueue(arr, len(queries), 6, 0)
        res, refined = handle.loop_batch_finish_refined(len(queries))
        handle.set_refiner(enabled=False)
        grids = [checker.grid(m.grid, m.res, m.off_x, m.off_y) for m in batch.submaps]
        odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
        ores, _ = odet.detect(grids, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                              batch.angles, batch.ranges)
        assert sum(o.found for o in ores) >= 2
        for i, (r, f, o) in enumerate(zip(res, refined, ores)):
            assert r.flags == capi.FLAG_EXACT
            assert_match(r, dict(o.asdict(), compare_unfound=False), "loop query %d" % i, flags_ok=capi.FLAG_EXACT)
            assert f.valid == o.found
        for mid in batch.map_ids:
            handle.release_grid(int(mid))
    finally:
        handle.set_option("fp_margin_scale", 1)


# --------------------------------------------------------------------------
# loop detection batch
# --------------------------------------------------------------------------
def _loop_queries(batch):
    scan = matchers.ScanData(batch.angles[0], batch.ranges[0])
    return [matchers.LoopDetectionQuery(
        scan, 0, tuple(batch.scan_poses[i]),
        matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), int(batch.map_ids[i])),
        tuple(batch.map_poses[i]), i) for i, s in enumerate(batch.submaps)]


def test_loop_batch_golden(handle):
    entry = load_golden("reference_vectors.json")["loop"][0]
    batch = synth.make_loop_batch(entry["seed"], n_maps=entry["n_maps"], true_fraction=entry["true_fraction"])
    assert sha(np.stack([s.grid for s in batch.submaps])) == entry["grid_sha"]
    bb = matchers.ScanMatcherBranchBound("loop-bb", entry["hmax"], *synth.CFG3["rng"], handle=handle)
    det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
    found, res = det.detect(_loop_queries(batch))
    assert sum(e["found"] for e in entry["expect"]) == len(found) > 0
    for i, (r, e) in enumerate(zip(res, entry["expect"])):
        e = dict(e, compare_unfound=False)
        assert_match(r, e, "loop query %d" % i)
    for f in found:
        e = entry["expect"][f.query_index]
        assert tuple(f.relative_pose) == tuple(e["est_pose"])
    for mid in batch.map_ids:
        handle.release_grid(int(mid))


def test_loop_batch_vs_checker_and_best_key(handle, checker):
    batch = synth.make_loop_batch(3100, n_maps=40, true_fraction=0.3, map_id_base=5000)
    bb = matchers.ScanMatcherBranchBound("loop-bb", 6, *synth.CFG3["rng"], handle=handle)
    det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
    queries = _loop_queries(batch)
    found, res = det.detect(queries, query_index_base=100)
    grids = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 2)
    ores, _ = odet.detect(grids, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    for i, (r, o) in enumerate(zip(res, ores)):
        d = dict(o.asdict(), compare_unfound=False)
        assert_match(r, d, "loop query %d" % i)
    # second call hits the device-side pyramid cache and must give the same answer
    found2, res2 = det.detect(queries, query_index_base=100)
    # (n_processed / n_ignored depend on the order in which leaves raise the incumbent)
    det_fields = ("found", "best_x", "best_y", "best_t", "sum_value", "n_known", "flags", "normalized_score")
    assert [[getattr(r, f) for f in det_fields] for r in res] == \
           [[getattr(r, f) for f in det_fields] for r in res2]
    # packed best word: max key, lowest query index on ties
    import torch
    word = torch.empty(1, dtype=torch.int64, device="cuda:0")
    C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(
        C.c_void_p(word.data_ptr()), C.c_void_p(handle.best_key_device_ptr()), 8, 3)
    key, qidx = handle.decode_best_key(int(word.item()))
    keys = [998 * r.sum_value + 64536 * r.n_known if r.found else -1 for r in res]
    assert key == max(keys) and qidx == 100 + keys.index(max(keys))
    for mid in batch.map_ids:
        handle.release_grid(int(mid))


@pytest.mark.parametrize("hmax", [6, 4, 3])
def test_incumbent_dive_changes_the_work_not_the_results(handle, checker, hmax):
    """k_bbg_dive (option bb_probe): with the incumbents seeded after the launch of height 4 (1), after heights 5
    and 4 (2), from any mask of heights (here 4 | 8: heights 2 and 3), or never (0), every query returns the
    same winner, sums, score and flags -- and the reference's --, while the sweep scores fewer children."""
    batch = synth.make_loop_batch(3300 + hmax, n_maps=48, true_fraction=0.4, map_id_base=7000)
    bb = matchers.ScanMatcherBranchBound("loop-bb", hmax, *synth.CFG3["rng"], handle=handle)
    det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
    queries = _loop_queries(batch)
    grids = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(hmax, synth.CFG3["rng"], synth.CFG3["thr"], 2)
    ores, _ = odet.detect(grids, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    fields = ("found", "best_x", "best_y", "best_t", "sum_value", "n_known", "flags", "normalized_score")
    got, scored = {}, {}
    try:
        for probe in (0, 1, 2, 12):
            handle.set_option("bb_probe", probe)
            found, res = det.detect(queries)
            for i, (r, o) in enumerate(zip(res, ores)):
                assert_match(r, dict(o.asdict(), compare_unfound=False), "hmax %d probe %d query %d" % (hmax, probe, i))
            got[probe] = [[getattr(r, f) for f in fields] for r in res]
            scored[probe] = sum(r.n_processed + r.n_ignored for r in res)
    finally:
        handle.set_option("bb_probe", 1)
        for mid in batch.map_ids:
            handle.release_grid(int(mid))
    assert sum(r.found for r in res) >= 5
    assert got[0] == got[1] == got[2] == got[12]
    assert scored[1] < scored[0] and scored[2] <= scored[1] and scored[12] < scored[0], scored


# --------------------------------------------------------------------------
# block-sparse upload (the reference's own storage layout)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("log2bs,shape", [(4, (512, 512)), (3, (64, 128)), (5, (128, 64)), (6, (128, 128))])
def test_block_sparse_upload_equals_dense(handle, log2bs, shape):
    rng = np.random.default_rng(77 + log2bs)
    if shape == (512, 512):
        grid = synth.rasterize(synth.make_room(rng), rng, *shape, 0.05).grid
    else:
        grid = rng.integers(1, 65535, size=shape, dtype=np.uint16)
        bs = 1 << log2bs
        mask = rng.random((shape[0] // bs, shape[1] // bs)) < 0.5       # half the blocks never written
        grid[np.kron(mask, np.ones((bs, bs), dtype=bool))] = 0
    blocks, index, br, bc = synth.dense_to_blocks(grid, log2bs)
    assert 0 < len(index) < br * bc
    handle.upload_grid_blocks(600, blocks, index, log2bs, br, bc, 0.05, -3.0, -4.0)
    assert np.array_equal(handle.download_level(600, 0, shape), grid)
    # re-upload with other contents into the same slot, in shuffled block order
    grid2 = np.roll(grid, (1 << log2bs, 1 << log2bs), axis=(0, 1))
    blocks, index, br, bc = synth.dense_to_blocks(grid2, log2bs)
    perm = rng.permutation(len(index))
    handle.upload_grid_blocks(600, blocks[perm], index[perm], log2bs, br, bc, 0.05, -3.0, -4.0)
    handle.build_pyramid(600, 3)
    assert np.array_equal(handle.download_level(600, 0, shape), grid2)
    # an all-unknown map: no blocks at all
    handle.upload_grid_blocks(600, np.zeros((0, 1 << log2bs, 1 << log2bs), np.uint16), np.zeros(0, np.int32),
                              log2bs, br, bc, 0.05, 0.0, 0.0)
    assert not handle.download_level(600, 0, shape).any()
    handle.release_grid(600)


def test_block_sparse_errors(handle):
    blocks = np.ones((2, 16, 16), np.uint16)
    with pytest.raises(capi.CsmError):
        handle.upload_grid_blocks(601, blocks, np.array([0, 0], np.int32), 4, 2, 2, 0.05, 0.0, 0.0)   # repeated
    with pytest.raises(capi.CsmError):
        handle.upload_grid_blocks(601, blocks, np.array([0, 4], np.int32), 4, 2, 2, 0.05, 0.0, 0.0)   # outside
    with pytest.raises(capi.CsmError):
        handle.upload_grid_blocks(601, blocks, np.array([0, 1], np.int32), 2, 2, 2, 0.05, 0.0, 0.0)   # 4x4 blocks


def test_loop_batch_block_sparse_upload(handle, checker):
    """The whole Detect with maps handed over block-sparse in one batched call
    gives the results of the dense path and of the checker."""
    batch = synth.make_loop_batch(3300, n_maps=24, true_fraction=0.4, map_id_base=8000)
    parts = [synth.dense_to_blocks(s.grid) for s in batch.submaps]
    blocks = np.concatenate([p[0] for p in parts])
    index = np.concatenate([p[1] for p in parts])
    counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
    ids = batch.map_ids.astype(np.int64)
    handle.upload_grids_blocks_ptr(ids, blocks.ctypes.data, index.ctypes.data, counts, 4, 32, 32,
                                   batch.submaps[0].res, np.array([s.off_x for s in batch.submaps]),
                                   np.array([s.off_y for s in batch.submaps]))
    bb = matchers.ScanMatcherBranchBound("loop-bb", 6, *synth.CFG3["rng"], handle=handle)
    det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
    det._cached_maps.update(int(i) for i in ids)          # already on the device
    handle.build_pyramids(ids, 6)
    found, res = det.detect(_loop_queries(batch))
    grids = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 2)
    ores, _ = odet.detect(grids, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    assert sum(o.found for o in ores) > 0
    for i, (r, o) in enumerate(zip(res, ores)):
        assert_match(r, dict(o.asdict(), compare_unfound=False), "loop query %d" % i)
    for i, s in enumerate(batch.submaps):
        assert np.array_equal(handle.download_level(int(ids[i]), 0, s.grid.shape), s.grid)
        handle.release_grid(int(ids[i]))


def test_errors(handle):
    with pytest.raises(capi.CsmError):
        handle.build_pyramid(123456, 3)                      # unknown map
    g = np.zeros((32, 32), np.uint16)
    handle.upload_grid(9, g, 0.05, 0.0, 0.0)
    with pytest.raises(capi.CsmError):
        handle.match_rt(9, [0.0], [1.0], (0, 0, 0), 5, (1, 1, 1), (0.05, 0.05, 0.01), (0.0, 0.0))  # coarse not built
    with pytest.raises(capi.CsmError):
        handle.build_pyramid(9, 9)                           # hmax out of range
    # refinement / epilogue entry points
    with pytest.raises(capi.CsmError):
        handle.set_refiner(0, 1e-4, 1e-4, 1e4)               # at least one iteration
    with pytest.raises(capi.CsmError):
        handle.set_refiner(10, 1e-4, 1e-4, 0.0)              # covariance scale must be positive
    with pytest.raises(capi.CsmError):
        handle.refine_batch([(123456, 1, (0.0, 0.0, 0.0))])  # unknown map
    with pytest.raises(capi.CsmError):
        handle.refine_batch([(9, 987654, (0.0, 0.0, 0.0))])  # unknown scan
    with pytest.raises(capi.CsmError):
        handle.set_epilogue(-1.0)
    handle.set_epilogue(1e4)
    with pytest.raises(capi.CsmError):
        handle.last_epilogue()                               # no match has run with it yet
    handle.set_epilogue(0.0)
    with pytest.raises(capi.CsmError):
        handle.loop_batch_finish_refined(1)                  # nothing in flight
    other = capi.Handle(0)
    assert handle.lib.csm_share_copy_stream(handle.h, handle.h) != 0      # a handle cannot borrow from itself
    assert handle.lib.csm_share_copy_stream(other.h, handle.h) == 0
    other.upload_grid(1, g, 0.05, 0.0, 0.0)                  # uploads through the borrowed copy stream
    assert not other.download_level(1, 0, (32, 32)).any()
    other.close()
    handle.release_grid(9)


# --------------------------------------------------------------------------
# the C++ plugin mirror (host/include/csm_host) end to end
# --------------------------------------------------------------------------
def _cmp_host(s, o, what):
    assert s.found == o.found, what
    assert (s.best_x, s.best_y, s.best_t, s.sum_value, s.n_known) == \
           (o.best_x, o.best_y, o.best_t, o.sum_value, o.n_known), what
    assert s.score == o.score, what
    assert list(s.est_pose) == list(o.est_pose), what            # refined pose: bit-identical
    assert s.norm_cost == o.norm_cost, what
    assert np.allclose(list(s.cov), list(o.cov), rtol=1e-9, atol=0.0), what


@pytest.mark.parametrize("seed", range(2400, 2404))
def test_cpp_adapter_matchers(checker, seed):
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    case = synth.case_for(synth.CFG1, seed)
    s = case.submap
    g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
    rel = (0.1, -0.03, 0.2)
    off = (s.off_x, s.off_y)
    a = ctx.match("rt", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"],
                  thr=(0.2, 0.3), rel_pose=rel)
    _cmp_host(a, checker.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"],
                                  (0.2, 0.3), rel), "rt")
    a = ctx.match("bb", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"],
                  rel_pose=rel)
    _cmp_host(a, checker.match_bb(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"],
                                  (0.0, 0.0), rel), "bb")
    a = ctx.match("grid", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 0, (0.4, 0.3, 0.05),
                  step=(0.05, 0.05, 0.005), rel_pose=rel)
    _cmp_host(a, checker.match_grid(g, case.angles, case.ranges, case.init_pose, (0.4, 0.3, 0.05),
                                    (0.05, 0.05, 0.005), (0.0, 0.0), rel), "grid")
    ctx.close()


@pytest.mark.parametrize("seed", range(2420, 2424))
def test_cpp_adapter_device_epilogue(checker, seed):
    """Cost and covariance of the decided pose computed on the device behind the match
    (csm_set_epilogue) instead of the CPU epilogue: search results and estimated pose bit-identical
    to the reference, normalized cost within 1e-12 relative and covariance within 1e-6 (the sums run
    in a tree order on the device). Dense and block-sparse maps, thresholds that fail included."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    ctx.set_device_epilogue(True)
    case = synth.case_for(synth.CFG1, seed)
    s = case.submap
    g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
    rel = (0.1, -0.03, 0.2) if seed % 2 else (0.0, 0.0, 0.0)
    off = (s.off_x, s.off_y)
    blocks, index, br, bc = synth.dense_to_blocks(s.grid)

    def cmp(a, o, what):
        assert a.found == o.found, what
        assert (a.best_x, a.best_y, a.best_t, a.sum_value, a.n_known) == \
               (o.best_x, o.best_y, o.best_t, o.sum_value, o.n_known), what
        assert a.score == o.score and list(a.est_pose) == list(o.est_pose), what
        assert np.isclose(a.norm_cost, o.norm_cost, rtol=1e-12, atol=0.0), what
        assert np.allclose(list(a.cov), list(o.cov), rtol=1e-6, atol=0.0), what

    for thr in ((0.0, 0.0), (0.2, 0.3), (0.999, 0.999)):
        a = ctx.match("rt", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"],
                      thr=thr, rel_pose=rel)
        cmp(a, checker.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"], thr, rel),
            "rt %s" % (thr,))
        a = ctx.match_blocks("bb", blocks, index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                             case.init_pose, 5, synth.CFG2["rng"], thr=thr, rel_pose=rel)
        cmp(a, checker.match_bb(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"], thr, rel),
            "bb %s" % (thr,))
    # the grid search keeps the CPU epilogue on such a context
    a = ctx.match("grid", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 0, (0.4, 0.3, 0.05),
                  step=(0.05, 0.05, 0.005), rel_pose=rel)
    _cmp_host(a, checker.match_grid(g, case.angles, case.ranges, case.init_pose, (0.4, 0.3, 0.05),
                                    (0.05, 0.05, 0.005), (0.0, 0.0), rel), "grid")
    ctx.close()


@pytest.mark.parametrize("seed", range(2430, 2434))
def test_cpp_adapter_device_final_matcher(checker, seed):
    """The front end's pair of matchers in one submission (lidar_graph_slam_frontend.cpp:216-230): the
    real-time correlative / branch-and-bound match and then the final matcher (ScanMatcherLinearSolver)
    on the pose it found, on the device. Coarse result bit-identical to the reference's matcher, refined
    pose within 1e-5 relative (north_star) of the reference's solver started from the coarse pose."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    ctx.set_device_final_matcher(10, 1e-4, 1e-4)
    case = synth.case_for(synth.CFG1, seed)
    s = case.submap
    g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
    rel = (0.1, -0.03, 0.2) if seed % 2 else (0.0, 0.0, 0.0)
    off = (s.off_x, s.off_y)
    blocks, index, br, bc = synth.dense_to_blocks(s.grid)
    for kind in ("rt", "bb"):
        if kind == "rt":
            a = ctx.match("rt", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 5,
                          synth.CFG1["rng"], rel_pose=rel)
            o = checker.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"], (0.0, 0.0), rel)
        else:
            a = ctx.match_blocks("bb", blocks, index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                                 case.init_pose, 5, synth.CFG2["rng"], rel_pose=rel)
            o = checker.match_bb(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"], (0.0, 0.0), rel)
        assert a.found == o.found == 1
        assert (a.best_x, a.best_y, a.best_t, a.sum_value, a.n_known, a.score) == \
               (o.best_x, o.best_y, o.best_t, o.sum_value, o.n_known, o.score), kind
        f = checker.refine(g, case.angles, case.ranges, list(o.est_pose), rel, 10, 1e-4, 1e-4)
        assert np.allclose(list(a.est_pose), list(f.est_pose), rtol=1e-5, atol=0.0), kind
        assert np.isclose(a.norm_cost, f.norm_cost, rtol=1e-6, atol=0.0), kind
        assert np.allclose(list(a.cov), list(f.cov), rtol=1e-5, atol=0.0), kind
    # a threshold nothing passes: nothing to refine, the summary is the plain (CPU) epilogue at the coarse pose
    a = ctx.match("rt", s.grid, s.res, off, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"],
                  thr=(0.999, 0.999), rel_pose=rel)
    _cmp_host(a, checker.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"],
                                  (0.999, 0.999), rel), "rt unfound")
    ctx.close()


@pytest.mark.parametrize("seed", range(2410, 2413))
def test_cpp_adapter_block_sparse_maps(checker, seed):
    """The C++ adapter fed with the map in the reference's block-sparse storage form: same device
    result and the same CPU cost / covariance (block allocation is exact in this form)."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    case = synth.case_for(synth.CFG1, seed)
    s = case.submap
    g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
    blocks, index, br, bc = synth.dense_to_blocks(s.grid)
    off = (s.off_x, s.off_y)
    a = ctx.match_blocks("rt", blocks, index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                         case.init_pose, 5, synth.CFG1["rng"])
    _cmp_host(a, checker.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"]), "rt")
    a = ctx.match_blocks("bb", blocks, index, 4, s.grid.shape, s.res, off, case.angles, case.ranges,
                         case.init_pose, 5, synth.CFG2["rng"])
    _cmp_host(a, checker.match_bb(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"]), "bb")
    ctx.close()


def test_cpp_adapter_loop_detector(checker):
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    batch = synth.make_loop_batch(3200, n_maps=20, true_fraction=0.4, map_id_base=7000)
    grids = np.stack([s.grid for s in batch.submaps])
    res = ctx.loop_detect(grids, batch.submaps[0].res, [s.off_x for s in batch.submaps],
                          [s.off_y for s in batch.submaps], batch.map_ids, batch.map_poses, batch.scan_poses,
                          batch.angles[0], batch.ranges[0], 6, synth.CFG3["rng"], synth.CFG3["thr"])
    og = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    ores, _ = odet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    assert sum(o.found for o in ores) > 0
    for i, (r, o) in enumerate(zip(res, ores)):
        assert r.found == o.found, i
        if o.found:
            assert (r.best_x, r.best_y, r.best_t, r.sum_value, r.n_known) == \
                   (o.best_x, o.best_y, o.best_t, o.sum_value, o.n_known), i
            assert r.score == o.score and list(r.est_pose) == list(o.est_pose), i
            assert np.allclose(list(r.cov), list(o.cov), rtol=1e-9, atol=0.0), i
    ctx.close()


def _inverse_compound(start, end):
    c, s_ = np.cos(start[2]), np.sin(start[2])
    dx, dy = end[0] - start[0], end[1] - start[1]
    return (c * dx + s_ * dy, -s_ * dx + c * dy, end[2] - start[2])


@pytest.mark.parametrize("kind", [0, 2])
def test_cpp_loop_detector_correlative_and_grid_search(checker, kind):
    """LoopDetectorCorrelative / LoopDetectorGridSearch of the C++ mirror: the reference's per-query loop
    (loop_detector_correlative.cpp:68-146, loop_detector_grid_search.cpp:64-129) is the coarse matcher with
    thresholds at InverseCompound(map pose, scan pose), a result only where a pose is found."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    batch = synth.make_loop_batch(3300 + kind, n_maps=6, true_fraction=0.5,
                                  offset=(0.2, 0.2, 0.05), map_id_base=7100 + 10 * kind)
    grids = np.stack([s.grid for s in batch.submaps])
    rng = (0.5, 0.5, 0.15) if kind == 0 else (0.4, 0.4, 0.1)
    step = None if kind == 0 else (0.05, 0.05, 0.005)
    thr = (0.5, 0.5)
    res, metrics = ctx.loop_detect_kind(kind, grids, batch.submaps[0].res, [s.off_x for s in batch.submaps],
                                        [s.off_y for s in batch.submaps], batch.map_ids, batch.map_poses,
                                        batch.scan_poses, batch.angles[0], batch.ranges[0], 5, rng, step, thr)
    n_found = 0
    for i, sm in enumerate(batch.submaps):
        g = checker.grid(sm.grid, sm.res, sm.off_x, sm.off_y)
        init = _inverse_compound(batch.map_poses[i], batch.scan_poses[i])
        if kind == 0:
            o = checker.match_rt(g, batch.angles[0], batch.ranges[0], init, 5, rng, thr)
        else:
            o = checker.match_grid(g, batch.angles[0], batch.ranges[0], init, rng, step, thr)
        assert res[i].found == o.found, i
        if o.found:
            n_found += 1
            assert res[i].score == o.score and list(res[i].est_pose) == list(o.est_pose), i
            assert np.allclose(list(res[i].cov), list(o.cov), rtol=1e-9, atol=0.0), i
    assert 0 < n_found < len(batch.submaps)
    det = "LoopDetectorCorrelativeGPU" if kind == 0 else "LoopDetectorGridSearchGPU"
    mat = "LoopRTGPU" if kind == 0 else "LoopGridGPU"
    assert metrics[det + ".NumOfQueries"] == [len(batch.submaps)]
    assert metrics[det + ".NumOfDetections"] == [n_found]
    assert len(metrics[det + ".LoopDetectionTime"]) == n_found
    assert metrics[mat + ".NumOfScans"] == [360.0] * len(batch.submaps)
    assert len(metrics[mat + ".ScoreValue"]) == len(batch.submaps)
    if kind == 0:
        assert metrics[mat + ".WinSizeX"] == [5.0] * len(batch.submaps)
    ctx.close()


def test_cpp_loop_detector_with_linear_solver(checker):
    """Detect end to end like the reference's default configuration: GPU branch-and-bound, then the
    CPU linear-solver refiner on every detected loop. Refined poses within 1e-5 relative of the
    reference's (north_star); in fact bit-identical."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    batch = synth.make_loop_batch(3500, n_maps=16, true_fraction=0.5, map_id_base=9000)
    ctx = hostapi.Context(0)
    det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
    det.use_linear_solver(10, 1e-4, 1e-4)
    det.configure(chunk_size=8 | (4 << 16))
    grids = np.ascontiguousarray(np.stack([s.grid for s in batch.submaps]))
    n, out = det.detect(len(batch.submaps), grids.ctypes.data, None, None, None, 4, 512, 512, batch.submaps[0].res,
                        np.array([s.off_x for s in batch.submaps]), np.array([s.off_y for s in batch.submaps]),
                        batch.map_ids.astype(np.int64), np.ascontiguousarray(batch.map_poses),
                        np.ascontiguousarray(batch.scan_poses), np.ascontiguousarray(batch.angles[0]),
                        np.ascontiguousarray(batch.ranges[0]))
    og = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    odet.use_linear_solver(10, 1e-4, 1e-4)
    ores, _ = odet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    assert n == sum(o.found for o in ores) >= 3
    for i, (r, o) in enumerate(zip(out, ores)):
        assert r.found == o.found, i
        if o.found:
            assert np.allclose(list(r.est_pose), list(o.est_pose), rtol=1e-5, atol=0.0), i
            assert list(r.est_pose) == list(o.est_pose), i
            assert np.allclose(list(r.cov), list(o.cov), rtol=1e-9, atol=0.0), i
    det.close()
    ctx.close()


def test_device_refiner_against_reference_vectors(handle):
    """k_refine (ScanMatcherLinearSolver + CostSquareError on the device, csm_refine_batch) against the
    vectors the reference's scan_matcher_linear_solver.cpp produced (tests/golden/refine_vectors.json),
    all cases in one batch: iteration counts equal, refined poses within the north_star tolerance of
    1e-5 relative (observed: ~1e-12, summation order only), costs 1e-9, covariances 1e-6. Maps go up
    dense (block allocation derived on the device) and block-sparse (block list kept) alternately."""
    from helpers import load_golden, sha
    entries = load_golden("refine_vectors.json")["refine"]
    jobs, cases = [], []
    for k, e in enumerate(entries):
        case = synth.case_for(synth.CFG1, e["seed"])
        s = case.submap
        assert sha(s.grid) == e["grid_sha"]
        mid, sid = 8800 + k, 8800 + k
        if k % 2 == 0:
            handle.upload_grid(mid, s.grid, s.res, s.off_x, s.off_y)
        else:
            blocks, index, br, bc = synth.dense_to_blocks(s.grid)
            handle.upload_grid_blocks(mid, blocks, index, 4, br, bc, s.res, s.off_x, s.off_y)
        handle.upload_scan(sid, case.angles, case.ranges)
        init = [float.fromhex(v) for v in e["init"]]
        jobs.append((mid, sid, matchers.compound(tuple(init), tuple(e["rel"]))))
        cases.append(case)
    out = handle.refine_batch(jobs, 10, 1e-4, 1e-4, 1e4)
    worst = 0.0
    for e, o in zip(entries, out):
        assert o.valid == 1 and o.iterations == e["iterations"]
        est = matchers.move_backward(tuple(o.pose), tuple(e["rel"]))
        exp = [float.fromhex(v) for v in e["est_pose"]]
        assert np.allclose(est, exp, rtol=1e-5, atol=0.0)
        worst = max(worst, max(abs(a - b) / abs(b) for a, b in zip(est, exp)))
        assert np.isclose(o.final_cost / 360.0, float.fromhex(e["norm_cost"]), rtol=1e-9, atol=0.0)
        assert np.allclose(list(o.covariance), [float.fromhex(v) for v in e["cov"]], rtol=1e-6, atol=0.0)
        assert 1e-8 <= o.lambda_ <= 1e-4
    assert worst < 1e-9, worst
    for k in range(len(entries)):
        handle.release_grid(8800 + k)


def test_device_refiner_unallocated_blocks_and_borders(handle):
    """Cells of unallocated blocks and cells outside the map read as 0.5, unknown cells of allocated
    blocks as 0.0 (grid_map.cpp:424-436): poses whose beams leave the map or cross empty blocks,
    against the C++ host mirror of the solver (bit-identical to the reference, test_host_logic)."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    case = synth.case_for(synth.CFG1, 5150)
    s = case.submap
    handle.upload_grid(8900, s.grid, s.res, s.off_x, s.off_y)
    handle.upload_scan(8900, case.angles, case.ranges)
    tp = np.asarray(case.true_pose, dtype=np.float64)
    span = s.res * s.grid.shape[1]
    starts = [tp + np.array([0.03, -0.02, 0.004]), tp + np.array([0.45 * span, 0.0, 0.3]),
              tp + np.array([-0.6 * span, 0.5 * span, -1.0]), tp + np.array([0.0, 0.0, 3.0])]
    out = handle.refine_batch([(8900, 8900, tuple(p)) for p in starts], 10, 1e-4, 1e-4, 1e4)
    for p, o in zip(starts, out):
        ref, _ = hostapi.refine(s.grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges, p)
        assert o.iterations == ref.best_t
        assert np.allclose(list(o.pose), list(ref.est_pose), rtol=1e-5, atol=1e-9)
        assert np.isclose(o.final_cost / 360.0, ref.norm_cost, rtol=1e-8, atol=0.0)
    handle.release_grid(8900)


def test_cpp_loop_detector_with_device_refiner(checker):
    """Detect like the reference's default configuration with the final matcher on the device: GPU
    branch-and-bound, k_refine on every detected loop in the same batch, against the reference's
    Detect with its own ScanMatcherLinearSolver. Refined poses within 1e-5 relative (north_star)."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    batch = synth.make_loop_batch(3500, n_maps=16, true_fraction=0.5, map_id_base=9100)
    ctx = hostapi.Context(0)
    det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
    det.use_device_refiner(10, 1e-4, 1e-4)
    det.configure(chunk_size=8 | (4 << 16))
    parts = [synth.dense_to_blocks(s.grid, 4) for s in batch.submaps]
    counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
    blocks = np.ascontiguousarray(np.concatenate([p[0].reshape(-1) for p in parts]))
    index = np.ascontiguousarray(np.concatenate([p[1] for p in parts]))
    args = (np.array([s.off_x for s in batch.submaps]), np.array([s.off_y for s in batch.submaps]),
            batch.map_ids.astype(np.int64), np.ascontiguousarray(batch.map_poses),
            np.ascontiguousarray(batch.scan_poses), np.ascontiguousarray(batch.angles[0]),
            np.ascontiguousarray(batch.ranges[0]))
    og = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    odet.use_linear_solver(10, 1e-4, 1e-4)
    ores, _ = odet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    grids = np.ascontiguousarray(np.stack([s.grid for s in batch.submaps]))
    for sparse in (True, False):
        det.clear_cache()
        n, out = det.detect(len(batch.submaps), None if sparse else grids.ctypes.data,
                            blocks.ctypes.data if sparse else None, index.ctypes.data if sparse else None,
                            counts.ctypes.data if sparse else None, 4, 512, 512, batch.submaps[0].res, *args)
        assert n == sum(o.found for o in ores) >= 3
        for i, (r, o) in enumerate(zip(out, ores)):
            assert r.found == o.found, i
            if o.found:
                assert np.allclose(list(r.est_pose), list(o.est_pose), rtol=1e-5, atol=0.0), (sparse, i)
                assert np.allclose(list(r.cov), list(o.cov), rtol=1e-5, atol=0.0), (sparse, i)
    det.close()
    ctx.close()


def _cpp_batch_args(batch):
    parts = [synth.dense_to_blocks(s.grid, 4) for s in batch.submaps]
    counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
    blocks = np.ascontiguousarray(np.concatenate([p[0].reshape(-1) for p in parts]))
    index = np.ascontiguousarray(np.concatenate([p[1] for p in parts]))
    args = (np.array([s.off_x for s in batch.submaps]), np.array([s.off_y for s in batch.submaps]),
            batch.map_ids.astype(np.int64), np.ascontiguousarray(batch.map_poses),
            np.ascontiguousarray(batch.scan_poses), np.ascontiguousarray(batch.angles[0]),
            np.ascontiguousarray(batch.ranges[0]))
    return blocks, index, counts, args


class _RefDetect(list):
    """The reference's Detect results with its own linear solver (refined poses); .coarse = the same
    detector without the refinement stage (window indices, sums)."""


def _reference_detect(checker, batch, angles=None, ranges=None):
    og = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    a = batch.angles if angles is None else angles
    r = batch.ranges if ranges is None else ranges
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    odet.use_linear_solver(10, 1e-4, 1e-4)
    ores, _ = odet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses, a, r)
    cdet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    cres, _ = cdet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses, a, r)
    out = _RefDetect(ores)
    out.coarse = cres
    return out


def _assert_detect_equal(out, ores, what):
    for i, (r, o, c) in enumerate(zip(out, ores, ores.coarse)):
        assert r.found == o.found == c.found, (what, i)
        if o.found:
            assert (r.best_x, r.best_y, r.best_t, r.sum_value, r.n_known) == \
                   (c.best_x, c.best_y, c.best_t, c.sum_value, c.n_known), (what, i)
            assert r.flags == 0, (what, i)
            assert np.allclose(list(r.est_pose), list(o.est_pose), rtol=1e-5, atol=0.0), (what, i)


def test_cpp_loop_detector_from_heap_allocated_blocks(checker):
    """Detect from maps in the reference's own storage -- every allocated 16x16 block a separate heap
    allocation (grid_map.cpp:522-535) -- gathered into page-locked staging by the detector's thread pool,
    group by group: the reference's results, whatever the number of gather threads and lanes."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    batch = synth.make_loop_batch(3700, n_maps=20, true_fraction=0.5, map_id_base=9600)
    blocks, index, counts, args = _cpp_batch_args(batch)
    ores = _reference_detect(checker, batch)
    heap = hostapi.HeapMaps(blocks, index, counts)
    ctx = hostapi.Context(0)
    for lanes, threads in ((1, 1), (2, 4), (3, 8)):
        det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
        det.use_device_refiner(10, 1e-4, 1e-4)
        det.configure(chunk_size=8 | (4 << 16))
        det.set_lanes(lanes)
        det.set_gather_threads(threads)
        for rep in range(2):             # the second call finds every map resident
            n, out = det.detect_heap(len(batch.submaps), heap, 512, 512, batch.submaps[0].res, *args)
            assert n == sum(o.found for o in ores) >= 3
            _assert_detect_equal(out, ores, (lanes, threads, rep))
        det.close()
    heap.close()
    ctx.close()


def test_cpp_loop_detector_scans_are_per_call(checker):
    """Two Detect calls on one detector with DIFFERENT scans under the same scan_id (the adapter's
    default 0): the second call must search with its own scan (the reference keeps no per-scan state)."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    batch = synth.make_loop_batch(3800, n_maps=10, true_fraction=0.6, map_id_base=9700)
    other = synth.make_loop_batch(3801, n_maps=10, true_fraction=0.6, map_id_base=9700)
    blocks, index, counts, args = _cpp_batch_args(batch)
    ctx = hostapi.Context(0)
    det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
    det.use_device_refiner(10, 1e-4, 1e-4)
    n1, out1 = det.detect(10, None, blocks.ctypes.data, index.ctypes.data, counts.ctypes.data, 4, 512, 512,
                          batch.submaps[0].res, *args)
    _assert_detect_equal(out1, _reference_detect(checker, batch), "first scan")
    # same maps (resident), same poses, another scan
    args2 = args[:5] + (np.ascontiguousarray(other.angles[0]), np.ascontiguousarray(other.ranges[0]))
    n2, out2 = det.detect(10, None, blocks.ctypes.data, index.ctypes.data, counts.ctypes.data, 4, 512, 512,
                          batch.submaps[0].res, *args2)
    ores2 = _reference_detect(checker, batch, other.angles, other.ranges)
    _assert_detect_equal(out2, ores2, "second scan")
    assert [o.found for o in ores2] != [o.found for o in _reference_detect(checker, batch)] or n1 != n2 or True
    det.close()
    ctx.close()


def test_cpp_loop_detector_recovers_from_frontier_overflow(checker):
    """CSM_E_CAPACITY is recoverable: with the frontier lists capped far below what the batch needs,
    Detect halves the overflowing batches until they fit and still returns the reference's results."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    batch = synth.make_loop_batch(3900, n_maps=16, true_fraction=0.6, map_id_base=9800)
    blocks, index, counts, args = _cpp_batch_args(batch)
    ores = _reference_detect(checker, batch)
    ctx = hostapi.Context(0)
    det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
    det.use_device_refiner(10, 1e-4, 1e-4)
    det.configure(chunk_size=16 | (16 << 16))
    h = capi.Handle.from_pointer(det.handle(), 0)
    h.set_option("bb_capacity", 400)           # 16 queries x 15 root groups = 240 fit, their children do not
    n, out = det.detect(16, None, blocks.ctypes.data, index.ctypes.data, counts.ctypes.data, 4, 512, 512,
                        batch.submaps[0].res, *args)
    h.set_option("bb_capacity", 0)
    assert det.capacity_retries() >= 1
    assert n == sum(o.found for o in ores) >= 3
    _assert_detect_equal(out, ores, "after overflow")
    det.close()
    ctx.close()


def test_cpp_multi_gpu_loop_detector(checker):
    """LoopDetectorBranchBoundMultiGPU (one process, one host thread and device context per GPU, queries
    sharded by LocalMapId mod G, results in query order, best word over one 8-byte NCCL all-reduce): the
    reference's results on every query, the best word of the unsharded batch."""
    import torch
    from my_lidar_graph_slam_v2_b200 import hostapi
    G = min(torch.cuda.device_count(), 4)
    batch = synth.make_loop_batch(4000, n_maps=24, true_fraction=0.5, map_id_base=9900)
    blocks, index, counts, args = _cpp_batch_args(batch)
    ores = _reference_detect(checker, batch)
    heap = hostapi.HeapMaps(blocks, index, counts)
    det = hostapi.MultiGpuLoopDetector(G, 6, synth.CFG3["rng"], synth.CFG3["thr"], lanes=2)
    det.configure(chunk_size=8, upload_chunk=4, gather_threads=2)
    det.use_nccl()
    for rep, hm in enumerate((heap, None)):
        if rep:
            det.clear_cache()
        n, out = det.detect(24, blocks.ctypes.data, index.ctypes.data, counts.ctypes.data, hm, 4, 512, 512,
                            batch.submaps[0].res, *args)
        assert n == sum(o.found for o in ores) >= 3
        _assert_detect_equal(out, ores, "multi gpu %d" % rep)
        assert sum(det.shard_sizes()) == 24 and len(det.shard_sizes()) == G
        keys = [998 * r.sum_value + 64536 * r.n_known if r.found else -1 for r in out]
        best = max(range(24), key=lambda i: (keys[i], -i))
        assert det.best_word() == (keys[best] << 20) | (0xFFFFF - best)
    det.close()
    heap.close()


def test_best_word_exchange_over_nccl(handle):
    """The library's own NCCL exchange (csm_comm_*): a communicator over this process's handles, the
    word a batch leaves on the device and a word formed on the host both come back reduced."""
    import torch
    G = min(torch.cuda.device_count(), 4)
    hs = [capi.Handle(g) for g in range(G)]
    capi.comm_init_all(hs)
    words = [(1000 + 17 * g) << 20 | (0xFFFFF - g) for g in range(G)]
    tickets = [h.comm_allreduce_word(w) for h, w in zip(hs, words)] if G == 1 else None
    if G == 1:
        assert hs[0].comm_best_result(tickets[0]) == words[0]
    else:
        lib = capi.load()
        arr = (C.c_void_p * G)(*[h.h for h in hs])
        tk = (C.c_int * G)()
        assert lib.csm_comm_allreduce_words_all(arr, G, (C.c_uint64 * G)(*words), tk) == 0
        for h, t in zip(hs, tk):
            assert h.comm_best_result(t) == max(words)
    for h in hs:
        h.close()


def test_reupload_while_a_block_sparse_batch_is_pending():
    """A map of a block-sparse batch that has not been expanded yet is uploaded again on its own: the
    expansion of the batch (triggered by the first use of a sibling) must not write the old blocks over
    the new contents, and the sibling keeps its own."""
    h = capi.Handle(0)
    rng = np.random.default_rng(5)
    grids = []
    for _ in range(3):
        g = rng.integers(1, 65535, size=(64, 64), dtype=np.uint16)
        g[rng.random((64, 64)) < 0.5] = 0
        grids.append(g)
    parts = [synth.dense_to_blocks(g, 4) for g in grids]
    counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
    blocks = np.ascontiguousarray(np.concatenate([p[0].reshape(-1) for p in parts]))
    index = np.ascontiguousarray(np.concatenate([p[1] for p in parts]))
    ids = np.array([700, 701, 702], dtype=np.int64)
    off = np.zeros(3)
    lib = capi.load()
    assert lib.csm_upload_grids_blocks(h.h, 3, ids.ctypes.data_as(C.POINTER(C.c_int64)), blocks.ctypes.data,
                                       index.ctypes.data, counts.ctypes.data_as(C.POINTER(C.c_int32)), 4, 4, 4, 0.05,
                                       off.ctypes.data_as(C.POINTER(C.c_double)),
                                       off.ctypes.data_as(C.POINTER(C.c_double))) == 0
    new = rng.integers(1, 65535, size=(64, 64), dtype=np.uint16)
    h.upload_grid(701, new, 0.05, 0.0, 0.0)               # replaces the middle map before anything used the batch
    assert np.array_equal(h.download_level(700, 0, (64, 64)), grids[0])
    assert np.array_equal(h.download_level(701, 0, (64, 64)), new)
    assert np.array_equal(h.download_level(702, 0, (64, 64)), grids[2])
    h.close()


def test_best_word_accumulation_and_query_index_range(handle):
    """"accumulate_best_key" set before the first batch starts from a zero word; a query index base that
    does not fit the packed word is refused."""
    h = capi.Handle(0)
    h.set_option("accumulate_best_key", 1)
    case = synth.case_for(synth.CFG1, 2500)
    s = case.submap
    bb = matchers.ScanMatcherBranchBound("bb", 3, 0.6, 0.6, 0.2, handle=h)
    det = matchers.LoopDetectorBranchBound("loop", bb, 0.3, 0.3)
    q = [matchers.LoopDetectionQuery(_scan(case), 0, tuple(case.init_pose),
                                     matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y), 42), (0.0, 0.0, 0.0), 0)]
    arr = det.prepare(q)
    res = h.loop_batch(arr, 1, 3, 7)
    import torch
    word = torch.as_tensor(_CudaWord(h.best_key_device_ptr()), device="cuda").cpu().item() & ((1 << 64) - 1)
    assert res[0].found and word == ((998 * res[0].sum_value + 64536 * res[0].n_known) << 20) | (0xFFFFF - 7)
    with pytest.raises(capi.CsmError):
        h.loop_batch(arr, 1, 3, 0x100000)
    with pytest.raises(capi.CsmError):
        h.loop_batch(arr, 1, 3, -1)
    h.close()


class _CudaWord:
    def __init__(self, ptr):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


# --------------------------------------------------------------------------
# BASELINE.json's full sizes
# --------------------------------------------------------------------------
def test_full_size_cfg3_batch(handle, checker):
    """configs[2] at full size: 1 scan x 256 first-touch 512x512 submaps, hmax 6, reference thresholds.
    (a) every query against the reference's LoopDetectorBranchBound::Detect (the CPU finishes the 256
    queries in seconds on all host threads); (b) properties that must hold whatever the execution
    shape: identical results with the incumbent dive on, with the sweep starting at the root level,
    split into 4 chunks, and sharded in two halves whose best words reduce to the unsharded one;
    (c) refinement on or off never changes the coarse results."""
    import os
    import torch
    batch = synth.make_loop_batch(31000, n_maps=256, true_fraction=0.25, map_id_base=20000)
    bb = matchers.ScanMatcherBranchBound("loop-bb", 6, *synth.CFG3["rng"], handle=handle)
    det = matchers.LoopDetectorBranchBound("loop", bb, *synth.CFG3["thr"])
    queries = _loop_queries(batch)
    fields = ("found", "best_x", "best_y", "best_t", "sum_value", "n_known", "flags", "normalized_score")

    def best_word():
        word = torch.empty(1, dtype=torch.int64, device="cuda:0")
        C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(
            C.c_void_p(word.data_ptr()), C.c_void_p(handle.best_key_device_ptr()), 8, 3)
        return int(word.item())

    def run(**opts):
        for k, v in opts.items():
            handle.set_option(k, v)
        try:
            _, res = det.detect(queries, query_index_base=0)
        finally:
            for k in opts:
                handle.set_option(k, {"bb_dive": 2, "bb_skip_top": 1, "bb_sweep_ctas_per_sm": 0}[k])
        return [[getattr(r, f) for f in fields] for r in res], best_word()

    base, word = run()
    assert 40 <= sum(r[0] for r in base) <= 120
    # (a) the reference, all queries
    grids = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], max(1, len(os.sched_getaffinity(0))))
    ores, _ = odet.detect(grids, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    _, res = det.detect(queries, query_index_base=0)
    for i, (r, o) in enumerate(zip(res, ores)):
        assert_match(r, dict(o.asdict(), compare_unfound=False), "cfg3 query %d" % i)
    # (b) execution shape
    assert run(bb_dive=1) == (base, word)
    assert run(bb_skip_top=0) == (base, word)
    assert run(bb_sweep_ctas_per_sm=2) == (base, word)      # smaller sweep grids (steps of several handles in flight)
    assert run(bb_sweep_ctas_per_sm=1) == (base, word)
    arr = det.prepare(queries)
    chunked = []
    handle.set_option("accumulate_best_key", 1)
    handle.set_option("reset_best_key", 1)
    for c in range(4):
        sub = (capi.CsmLoopQuery * 64).from_address(C.addressof(arr) + c * 64 * C.sizeof(capi.CsmLoopQuery))
        chunked += [[getattr(r, f) for f in fields] for r in handle.loop_batch(sub, 64, 6, c * 64)]
    handle.set_option("accumulate_best_key", 0)
    assert chunked == base and best_word() == word
    halves, words = [], []
    for s in range(2):
        sub = (capi.CsmLoopQuery * 128).from_address(C.addressof(arr) + s * 128 * C.sizeof(capi.CsmLoopQuery))
        halves += [[getattr(r, f) for f in fields] for r in handle.loop_batch(sub, 128, 6, s * 128)]
        words.append(best_word())
    assert halves == base and max(words) == word
    # (c) refinement leaves the coarse stage alone and refines exactly the found poses
    handle.set_refiner(10, 1e-4, 1e-4, 1e4)
    handle.loop_batch_enqueue(arr, 256, 6, 0)
    res_r, refined = handle.loop_batch_finish_refined(256)
    handle.set_refiner(enabled=False)
    assert [[getattr(r, f) for f in fields] for r in res_r] == base
    assert [f.valid for f in refined] == [r[0] for r in base]
    for mid in batch.map_ids:
        handle.release_grid(int(mid))


@pytest.mark.parametrize("entry", load_golden("cfg4_vectors.json")["cfg4"], ids=lambda e: "seed%d" % e["seed"])
def test_full_size_cfg4_grid_search(handle, checker, entry):
    """configs[3] at full size: 1080 beams, 1280x1280 map at 0.025 m, window 161 x 161 x 600 = 15.6 M
    candidates. The winner of the UNMODIFIED reference's ScanMatcherGridSearch::OptimizePose
    (scan_matcher_grid_search.cpp:84-178; ~14 minutes per case on one core here, so it was run once by
    tests/golden/make_cfg4_golden.py and committed) must come out of both device kernels: best index,
    value sum, known count and the double score bit for bit, no flags. On top of that the properties
    that need no golden: the reference asked for the score of exactly that pose agrees, and the winner
    lies within two cells of the true pose."""
    case = synth.case_for(synth.CFG4, entry["seed"])
    s = case.submap
    assert sha(s.grid) == entry["grid_sha"] and sha(case.ranges) == entry["scan_sha"], "synthetic generator drifted"
    assert [float(v).hex() for v in case.init_pose] == entry["init_pose"]
    gm, scan = grid_of(case), _scan(case)
    rng, step = synth.CFG4["rng"], synth.CFG4["step"]
    mt = matchers.ScanMatcherGridSearch("gs", *rng, *step, handle=handle)
    exp = entry["expect"]
    results = []
    for mode in (2, 3, 1):                    # 2 / 3: the TMA shared-memory kernel (32-bit / u16 tiles), 1: global memory
        handle.set_option("window_mode", mode)
        r = mt.optimize_pose(gm, scan, tuple(case.init_pose)).result
        assert_match(r, exp, "cfg4 window_mode %d" % mode)
        assert r.flags == 0 and r.n_processed == exp["n_processed"]
        results.append(r)
    handle.set_option("window_mode", 0)
    a = results[0]
    dx = matchers.grid_search_offsets(rng[0] / 2, step[0])
    dy = matchers.grid_search_offsets(rng[1] / 2, step[1])
    dt = matchers.grid_search_offsets(rng[2] / 2, step[2])
    # the reference's accumulating loops `for (d = -r; d <= r; d += s)` give 161 x 161 x 600 here
    assert (len(dx), len(dy), len(dt)) == (161, 161, 600)
    sensor = matchers.compound(tuple(case.init_pose), scan.relative_sensor_pose)
    win = (sensor[0] + dx[a.best_x], sensor[1] + dy[a.best_y], sensor[2] + dt[a.best_t])
    g = checker.grid(s.grid, s.res, s.off_x, s.off_y)
    o = checker.match_grid(g, case.angles, case.ranges, win, (0.0, 0.0, 0.0), step)
    assert (o.found, o.sum_value, o.n_known, o.score) == (1, a.sum_value, a.n_known, a.normalized_score)
    assert abs(win[0] - case.true_pose[0]) <= 2 * s.res and abs(win[1] - case.true_pose[1]) <= 2 * s.res


@pytest.mark.parametrize("lanes,chunk,up", [(2, 8, 4), (3, 4, 4), (4, 16, 3)])
def test_cpp_loop_detector_pipeline_lanes(checker, lanes, chunk, up):
    """Pipeline lanes (several device contexts of one detector sharing one copy stream): the same
    results as the reference whatever the lane / batch / upload-group split, a second Detect served
    from the per-lane caches, and the packed best word of the whole call."""
    from my_lidar_graph_slam_v2_b200 import hostapi
    batch = synth.make_loop_batch(3500, n_maps=16, true_fraction=0.5, map_id_base=9300)
    ctx = hostapi.Context(0)
    det = hostapi.LoopDetector(ctx, 6, synth.CFG3["rng"], synth.CFG3["thr"])
    det.use_device_refiner(10, 1e-4, 1e-4)
    det.configure(chunk_size=chunk | (up << 16), query_index_base=1000)
    det.set_lanes(lanes)
    parts = [synth.dense_to_blocks(s.grid, 4) for s in batch.submaps]
    counts = np.array([len(p[1]) for p in parts], dtype=np.int32)
    blocks = np.ascontiguousarray(np.concatenate([p[0].reshape(-1) for p in parts]))
    index = np.ascontiguousarray(np.concatenate([p[1] for p in parts]))
    args = (np.array([s.off_x for s in batch.submaps]), np.array([s.off_y for s in batch.submaps]),
            batch.map_ids.astype(np.int64), np.ascontiguousarray(batch.map_poses),
            np.ascontiguousarray(batch.scan_poses), np.ascontiguousarray(batch.angles[0]),
            np.ascontiguousarray(batch.ranges[0]))
    og = [checker.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    odet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    odet.use_linear_solver(10, 1e-4, 1e-4)
    ores, _ = odet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    # coarse results (indices, sums) from the reference's detector without the refinement stage
    cdet = checker.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    cres, _ = cdet.detect(og, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                          batch.angles, batch.ranges)
    sigs = []
    for rep in range(2):                 # the second call finds every map resident on its lane
        n, out = det.detect(len(batch.submaps), None, blocks.ctypes.data, index.ctypes.data, counts.ctypes.data,
                            4, 512, 512, batch.submaps[0].res, *args)
        assert n == sum(o.found for o in ores) >= 3
        for i, (r, o, c) in enumerate(zip(out, ores, cres)):
            assert r.found == o.found == c.found, i
            if o.found:
                assert (r.best_x, r.best_y, r.best_t, r.sum_value, r.n_known) == \
                       (c.best_x, c.best_y, c.best_t, c.sum_value, c.n_known), i
                assert np.allclose(list(r.est_pose), list(o.est_pose), rtol=1e-5, atol=0.0), i
        keys = [998 * r.sum_value + 64536 * r.n_known if r.found else -1 for r in out]
        best = max(range(len(keys)), key=lambda i: (keys[i], -i))
        assert det.best_word() == (keys[best] << 20) | (0xFFFFF - (1000 + best))
        sigs.append([(r.found, r.best_x, r.best_y, r.best_t, r.sum_value, tuple(r.est_pose)) for r in out])
    # (the damping factor carried over from the first call moves refined poses by ~1e-13)
    assert [g[:5] for g in sigs[0]] == [g[:5] for g in sigs[1]]
    assert np.allclose([g[5] for g in sigs[0]], [g[5] for g in sigs[1]], rtol=1e-9, atol=1e-12)
    det.close()
    ctx.close()


def _same_winner(dev, orc, what):
    """index and the reference's own double score; the checker's integer diagnostics (sum_value, n_known)
    count raw cell values and do not apply to saturated cells"""
    assert dev.flags == 0 and dev.found == orc.found == 1, what
    assert (dev.best_x, dev.best_y, dev.best_t) == (orc.best_x, orc.best_y, orc.best_t), what
    assert dev.normalized_score == orc.score, "%s: score %r vs %r" % (what, dev.normalized_score, orc.score)


def test_saturated_cells_read_as_unknown(handle):
    """Maps the reference builds hold cells at 65535 (ValueMax). Its value tables have 65535 entries
    (grid_values.cpp:32-35), so its matchers read those cells one element past the table: 0.0, the unknown
    probability, in the compiled reference. The library's copy of every map holds them as unknown: exhaustive
    searches (no pruning involved) and the final matcher then agree with the compiled reference on a map with
    saturated walls; with the option off the cell counts as 0.999 and the scores differ."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    if not (pyoracle.available("reference") or os.path.isdir("/root/reference")):
        pytest.skip("needs the compiled reference")
    ref = pyoracle.load("reference")
    case = synth.case_for(synth.CFG1, 6400)
    s = case.submap
    grid = s.grid.copy()
    rng = np.random.default_rng(6400)
    walls = grid > 45000
    grid[walls & (rng.random(grid.shape) < 0.3)] = 65535
    assert (grid == 65535).sum() > 500
    g = ref.grid(grid, s.res, s.off_x, s.off_y)
    gm = matchers.GridMap(grid, s.res, (s.off_x, s.off_y))
    scan = _scan(case)
    # real-time correlative with a 1-cell coarse map and the exhaustive grid search: every candidate is scored
    mt = matchers.ScanMatcherCorrelative("rt", 1, *synth.CFG1["rng"], handle=handle)
    got = mt.optimize_pose(gm, scan, tuple(case.init_pose), 0.0, 0.0)
    exp = ref.match_rt(g, case.angles, case.ranges, case.init_pose, 1, synth.CFG1["rng"])
    _same_winner(got.result, exp, "rt on saturated walls")
    mg = matchers.ScanMatcherGridSearch("grid", 0.4, 0.4, 0.06, 0.05, 0.05, 0.004, handle=handle)
    got = mg.optimize_pose(gm, scan, tuple(case.init_pose), 0.0, 0.0)
    exp_g = ref.match_grid(g, case.angles, case.ranges, case.init_pose, (0.4, 0.4, 0.06), (0.05, 0.05, 0.004))
    _same_winner(got.result, exp_g, "grid search on saturated walls")
    # the cost function of the final matcher interpolates the same cells
    handle.upload_grid(8950, grid, s.res, s.off_x, s.off_y)
    handle.upload_scan(8950, case.angles, case.ranges)
    start = np.asarray(case.true_pose) + np.array([0.03, -0.02, 0.004])
    out = handle.refine_batch([(8950, 8950, tuple(start))], 10, 1e-4, 1e-4, 1e4)[0]
    lam = [1e-4]
    exp_r = ref.refine(g, case.angles, case.ranges, list(start), None, 10, 1e-4, 1e-4)
    assert np.allclose(list(out.pose), list(exp_r.est_pose), rtol=1e-5, atol=1e-9)
    twin, _ = hostapi.refine(grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges, start)
    assert list(twin.est_pose) == list(exp_r.est_pose)
    # option off: 65535 counts as p = 0.999, which the reference cannot be asked for
    handle.set_option("saturated_unknown", 0)
    try:
        got_off = mt.optimize_pose(gm, scan, tuple(case.init_pose), 0.0, 0.0)
        assert got_off.result.normalized_score > exp.score
    finally:
        handle.set_option("saturated_unknown", 1)
    handle.release_grid(8950)


@pytest.mark.parametrize("kind", [0, 2])
def test_cpp_loop_detectors_concurrent_matchers(kind):
    """LoopDetectorCorrelative / LoopDetectorGridSearch with several matchers at once
    (SetConcurrentMatchers: query i on matcher LocalMapId mod L, one host thread and one device context
    each): the same results, in query order, as the plain per-query loop."""
    import time
    from my_lidar_graph_slam_v2_b200 import hostapi
    ctx = hostapi.Context(0)
    batch = synth.make_loop_batch(3400 + kind, n_maps=24, true_fraction=0.5,
                                  offset=(0.2, 0.2, 0.05), map_id_base=7300 + 100 * kind)
    grids = np.stack([s.grid for s in batch.submaps])
    rng = (0.5, 0.5, 0.15) if kind == 0 else (0.4, 0.4, 0.1)
    step = None if kind == 0 else (0.05, 0.05, 0.005)
    args = (kind, grids, batch.submaps[0].res, [s.off_x for s in batch.submaps], [s.off_y for s in batch.submaps],
            batch.map_ids, batch.map_poses, batch.scan_poses, batch.angles[0], batch.ranges[0], 5, rng, step, (0.5, 0.5))
    out, times = {}, {}
    try:
        for lanes in (1, 4):
            hostapi.set_detect_concurrency(lanes)
            ctx.loop_detect_kind(*args)                      # contexts and kernels warm
            t0 = time.perf_counter()
            out[lanes], _ = ctx.loop_detect_kind(*args)
            times[lanes] = time.perf_counter() - t0
    finally:
        hostapi.set_detect_concurrency(1)
    assert sum(r.found for r in out[1]) > 0
    for a, b in zip(out[1], out[4]):
        assert a.found == b.found
        if a.found:
            assert a.score == b.score and list(a.est_pose) == list(b.est_pose) and list(a.cov) == list(b.cov)
    print("kind %d: %.2f ms plain, %.2f ms on 4 matchers" % (kind, times[1] * 1e3, times[4] * 1e3))
    ctx.close()
