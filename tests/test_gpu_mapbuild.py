"""Map construction on the device against the UNMODIFIED reference.

GridMapBuilderGPU (host/src/map_builder.cpp over csm_map_* of the C ABI) rebuilds the front end's latest map
per appended scan like the reference's GridMapBuilder::UpdateLatestMap (mapping/grid_map_builder.cpp:288-371,
ComputeBoundingBoxAndScanPointsMapLocal :375-438): bounding box and Resize on the host, ray casting with the
scaled Bresenham (bresenham.cpp), missed/hit updates through the binary Bayes tables (grid_binary_bayes.cpp)
on the device. Bar: every cell value, the geometry (rows, cols, offset), the map pose and the block
allocation identical after every scan; the match on the resident map identical to the match on the
reference's map."""
import numpy as np
import pytest

from my_lidar_graph_slam_v2_b200 import synth

pytestmark = pytest.mark.gpu


def _trajectory(seed, n_scans, n_beams=360, step=(0.06, 0.025, 0.012)):
    rng = np.random.default_rng(seed)
    room = synth.make_room(rng)
    pose = synth.random_pose_in_room(room, rng)
    out = []
    for k in range(n_scans):
        p = pose + np.array(step) * k
        a, r = synth.raycast(room, p, n_beams, 0.01, 11.4, rng)
        out.append((p, a, r))
    return room, out


def _compare(dev, ref, what):
    d_dense, d_alloc, d_off, d_pose, d_bs = dev
    r_dense, r_alloc, r_off, r_pose, r_bs = ref
    assert d_bs == r_bs, what
    assert d_dense.shape == r_dense.shape, "%s: shape %s vs %s" % (what, d_dense.shape, r_dense.shape)
    assert tuple(d_off) == tuple(r_off), "%s: offset %s vs %s" % (what, d_off, r_off)
    assert np.array_equal(d_pose, r_pose), "%s: map pose" % what
    assert np.array_equal(d_alloc, r_alloc), "%s: block allocation differs in %d blocks" % (
        what, int((d_alloc != r_alloc).sum()))
    bad = np.argwhere(d_dense != r_dense)
    assert len(bad) == 0, "%s: %d cells differ, first %s dev %d ref %d" % (
        what, len(bad), bad[0], d_dense[tuple(bad[0])], r_dense[tuple(bad[0])])


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("seed,n_beams,rel", [(7001, 360, (0.0, 0.0, 0.0)), (7002, 1080, (0.12, -0.04, 0.3))])
def test_latest_map_per_scan(seed, n_beams, rel, fast):
    """14 scans into a 10-scan window: the window slides, the map is rebuilt from scratch per scan. `fast`: the
    host rotates each scan's polar points instead of calling libm per beam and re-evaluates the beams whose
    floored coordinates sit inside the guard band (the default); the other way is the reference's arithmetic."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    _, traj = _trajectory(seed, 14, n_beams)
    ctx = hostapi.Context(0)
    mb = hostapi.MapBuilder(ctx)
    mb.set_fast_hit_points(fast)
    ob = ref.map_builder()
    for k, (p, a, r) in enumerate(traj):
        n = mb.append(p, a, r, rel)
        ob.append(p, a, r, rel)
        assert n > 0
        _compare(mb.latest(), ob.latest(), "seed %d scan %d" % (seed, k))
    mb.close()
    ctx.close()


def test_guard_band_path_re_evaluates_exactly():
    """With the guard band widened to half a cell every fast hit point counts as "near a boundary": the bounding
    box sends every point through the exact re-evaluation; with a band of 0.05 cells about a third of the beams
    are re-evaluated one by one. The map is the reference's either way."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    _, traj = _trajectory(7006, 6, 360)
    ctx = hostapi.Context(0)
    knob = hostapi.MapBuilder(ctx)
    try:
        for band in (0.5, 0.05):
            mb = hostapi.MapBuilder(ctx, scans_for_latest_map=3)
            ob = ref.map_builder(scans_for_latest_map=3)
            knob.set_guard_band(band)
            exact = []
            for k, (p, a, r) in enumerate(traj):
                n = mb.append(p, a, r)
                ob.append(p, a, r)
                _compare(mb.latest(), ob.latest(), "band %.2f scan %d" % (band, k))
                exact.append((mb.last_exact_beams(), n))
            if band == 0.5:
                assert all(e == 0 for e, _ in exact)          # nothing was "fast" any more when inserted
            else:
                assert any(0 < e < n for e, n in exact), exact
            mb.close()
    finally:
        knob.set_guard_band(1e-9)
        knob.close()
    ctx.close()


def test_saturation_and_range_limits():
    """The same pose 40 times drives free cells to ValueMin and occupied cells to ValueMax = 65535, where
    the reference's next update reads one entry past its 65535-entry odds table and drops the cell to
    ValueMin (grid_values.cpp:72-74, grid_binary_bayes.cpp:316; deterministic with glibc, see
    map_builder.cpp): the device tables reproduce that. Beams outside [usable_range_min, usable_range_max]
    and the scan's own [min_range, max_range] are dropped as the reference drops them."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    rng = np.random.default_rng(7003)
    room = synth.make_room(rng)
    p = synth.random_pose_in_room(room, rng)
    a, r = synth.raycast(room, p, 360, 0.0, 11.4, rng)
    r = r.copy()
    r[::17] = 0.001           # under usable_range_min
    r[5::23] = 9.5            # over the builder's usable_range_max below
    ctx = hostapi.Context(0)
    kw = dict(scans_for_latest_map=40, usable_range_min=0.05, usable_range_max=9.0)
    mb = hostapi.MapBuilder(ctx, **kw)
    ob = ref.map_builder(**kw)
    for k in range(40):
        mb.append(p, a, r, min_range=0.02, max_range=8.0 if k % 2 else 50.0)
        ob.append(p, a, r, min_range=0.02, max_range=8.0 if k % 2 else 50.0)
    dev, orc = mb.latest(), ob.latest()
    _compare(dev, orc, "saturation")
    assert dev[0].max() == orc[0].max() and dev[0][dev[0] > 0].min() == orc[0][orc[0] > 0].min()
    mb.close()
    ctx.close()


def test_other_resolution_and_patch():
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    _, traj = _trajectory(7004, 6, 720, step=(-0.11, 0.07, -0.03))
    ctx = hostapi.Context(0)
    kw = dict(resolution=0.025, patch_size=8, scans_for_latest_map=4, prob_hit=0.7, prob_miss=0.4)
    mb = hostapi.MapBuilder(ctx, **kw)
    ob = ref.map_builder(**kw)
    for k, (p, a, r) in enumerate(traj):
        mb.append(p, a, r)
        ob.append(p, a, r)
        _compare(mb.latest(), ob.latest(), "scan %d" % k)
    mb.close()
    ctx.close()


def test_match_on_resident_map():
    """The front end's step: RT correlative match of the next scan on the latest map that never left the
    device, against the reference matcher on the reference's own latest map."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    room, traj = _trajectory(7005, 11, 360)
    ctx = hostapi.Context(0)
    ctx.set_device_epilogue()          # the map is not on the host: cost and covariance come from the device
    mb = hostapi.MapBuilder(ctx)
    ob = ref.map_builder()
    for p, a, r in traj[:10]:
        mb.append(p, a, r)
        ob.append(p, a, r)
    p, a, r = traj[10]
    init = p + np.array([0.07, -0.05, 0.02])
    dense, _, off, _, _ = ob.latest()
    g = ref.grid(dense, 0.05, off[0], off[1])
    exp = ref.match_rt(g, a, r, init, 5, synth.CFG1["rng"])
    got = mb.match_rt(a, r, init, 5, synth.CFG1["rng"])
    assert got.found == exp.found
    assert (got.best_x, got.best_y, got.best_t) == (exp.best_x, exp.best_y, exp.best_t)
    assert got.score == exp.score
    assert np.array_equal(np.array(got.est_pose[:3]), np.array(exp.est_pose[:3]))
    assert abs(got.norm_cost - exp.norm_cost) <= 1e-5 * abs(exp.norm_cost)
    assert np.allclose(np.array(got.cov[:9]), np.array(exp.cov[:9]), rtol=1e-5, atol=1e-12)
    mb.close()
    ctx.close()
