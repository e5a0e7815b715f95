"""The full loop (BASELINE configs[4]): C++ SlamPipeline -- front end and back end of
the reference's LidarGraphSlam around the device matchers and device-built maps -- against the same loop on
the reference's own components (oracle/ref_wrapper.cpp: RefSlam; GridMapBuilder, ScanMatcherCorrelative,
ScanMatcherLinearSolver, LoopSearcherNearest, LoopDetectorBranchBound are the compiled reference).

With the CPU twins of the final matcher (host_final_matchers) every quantity is bit-identical: scan-node
poses, pose-graph edges, local-map geometry and cells, loop candidates and detected loops. With the final
matchers on the device the trajectory agrees to 1e-6 and the discrete structure is the same."""
import numpy as np
import pytest

from my_lidar_graph_slam_v2_b200 import slam_settings, synth

pytestmark = pytest.mark.gpu

SETTINGS = dict(update_travel_dist=0.05, local_map_travel_dist=1.5, loop_detection_threshold=1.0,
                searcher_travel_dist=3.0, searcher_node_dist=5.0, searcher_candidates=24)


def _trip(seed, n_scans, spacing=0.1):
    rng = np.random.default_rng(seed)
    world = synth.corridor_world(rng)
    return synth.corridor_trajectory(world, n_scans, spacing, rng)


def _run(slam, trip):
    used = slam.run(trip["angles"], trip["ranges"], trip["odom"], trip["stamps"], 0.01, 11.3, finish=True)
    return used


@pytest.mark.parametrize("seed", [8101])
def test_full_loop_bit_identical_with_host_final_matchers(seed):
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    trip = _trip(seed, 130)
    st = slam_settings.pack(host_final_matchers=1, **SETTINGS)
    rs = ref.slam(st)
    ctx = hostapi.Context(0)
    gs = hostapi.SlamPipeline(ctx, st)
    assert _run(gs, trip) == _run(rs, trip) == 130
    gc, rc = gs.counters(), rs.counters()
    for k in ("scans_processed", "backend_steps", "backend_steps_with_candidates", "loop_queries", "loops_detected",
              "degenerations"):
        assert gc[k] == rc[k], (k, gc[k], rc[k])
    assert gc["loop_queries"] > 0 and gc["loops_detected"] > 0, gc
    assert gc["optimizer_calls"] == gc["optimizations"] > 0          # the seam was crossed
    assert np.array_equal(gs.scan_nodes(), rs.scan_nodes())
    assert np.array_equal(gs.edges(), rs.edges())
    gm, rm = gs.local_maps(), rs.local_maps()
    assert np.array_equal(gm, rm)
    for i in range(len(gm)):
        gd, ga = gs.local_map_cells(i)
        rd, ra = rs.local_map_cells(i)
        assert np.array_equal(ga, ra), "local map %d: block allocation" % i
        assert np.array_equal(gd, rd), "local map %d: %d cells differ" % (i, int((gd != rd).sum()))
    gl, rl = gs.loops(), rs.loops()
    assert np.array_equal(gl[:, :5], rl[:, :5])
    gs.close()
    ctx.close()


def test_full_loop_on_the_device():
    """Final matchers on the device (the fast configuration): same discrete structure, poses to 1e-6."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    trip = _trip(8102, 100)
    rs = ref.slam(slam_settings.pack(host_final_matchers=1, **SETTINGS))
    ctx = hostapi.Context(0)
    gs = hostapi.SlamPipeline(ctx, slam_settings.pack(host_final_matchers=0, **SETTINGS))
    assert _run(gs, trip) == _run(rs, trip) == 100
    gn, rn = gs.scan_nodes(), rs.scan_nodes()
    assert gn.shape == rn.shape and np.array_equal(gn[:, 6], rn[:, 6])
    assert np.abs(gn[:, :6] - rn[:, :6]).max() < 1e-6
    gm, rm = gs.local_maps(), rs.local_maps()
    assert gm.shape == rm.shape and np.array_equal(gm[:, 3:8], rm[:, 3:8])
    gc, rc = gs.counters(), rs.counters()
    assert gc["loop_queries"] == rc["loop_queries"] and gc["loops_detected"] == rc["loops_detected"]
    gs.close()
    ctx.close()


def test_full_loop_from_a_carmen_log(tmp_path):
    """The launcher's way in and out (slam_launcher.cpp:262-283, 171-181): the run written as a Carmen log
    (ROBOTLASER1 records), read back by the C++ CarmenLogReader and fed to SlamPipeline::RunLog gives the
    same pose graph as the arrays fed directly, bit for bit; the metrics land in <output>.metric.json under
    the reference's ids."""
    import json
    from my_lidar_graph_slam_v2_b200 import hostapi
    trip = _trip(8103, 90)
    n_beams = trip["ranges"].shape[1]
    start, inc = -np.pi, 2.0 * np.pi / n_beams
    angles = start + inc * np.arange(n_beams)
    path = str(tmp_path / "corridor.log")
    hostapi.write_carmen_log(path, trip["ranges"], trip["odom"], trip["stamps"], start, inc, 11.3)
    st = slam_settings.pack(host_final_matchers=0, **SETTINGS)
    ctx = hostapi.Context(0)
    direct = hostapi.SlamPipeline(ctx, st)
    assert direct.run(angles, trip["ranges"], trip["odom"], trip["stamps"], 0.0, 11.3, finish=True) == 90
    log = hostapi.CarmenLog(path=path)
    assert len(log.records()) == 180
    from_log = hostapi.SlamPipeline(ctx, st)
    from_log.record_metrics()
    assert from_log.run_carmen(log) == 90
    assert np.array_equal(from_log.scan_nodes(), direct.scan_nodes())
    assert np.array_equal(from_log.edges(), direct.edges())
    assert np.array_equal(from_log.local_maps(), direct.local_maps())
    assert np.array_equal(from_log.loops(), direct.loops())
    c = from_log.counters()
    out = str(tmp_path / "result")
    from_log.save_metrics(out)
    vs = json.load(open(out + ".metric.json"))["ValueSequences"]
    assert int(vs["Frontend.ProcessScanTime"]["NumOfSamples"]) == 90
    assert int(vs["Frontend.ScanMatchingTime"]["NumOfSamples"]) == 89
    assert vs["Frontend.NumOfScans"]["Values"].split() == [str(n_beams)] * 90
    assert vs["Frontend.ProcessFrame"]["Values"].split() == [str(i) for i in range(90)]
    ends = sum(int(vs[k]["NumOfSamples"]) for k in vs if k.startswith("Backend.EndAt"))
    assert ends == int(vs["Backend.ProcessTime"]["NumOfSamples"]) == c["backend_steps"]
    assert int(vs["Backend.LoopDetectionTime"]["NumOfSamples"]) == c["backend_steps_with_candidates"]
    assert any(k.startswith("LocalSlam.ScanMatcherCorrelative.") for k in vs)
    assert any(k.startswith("LoopDetector.BranchBound.") for k in vs)
    for s in (direct, from_log):
        s.close()
    log.close()
    ctx.close()
