"""BASELINE configs[4]: the full SLAM loop on a synthetic closed trajectory (60 x 40 m corridor world, drifting
odometry), driven through the C++ SlamPipeline: every matcher and every map on the device, the pose-graph
optimiser behind its seam (identity). scripts/cfg5_full_loop.py and bench.py time it next to the same loop
on the compiled reference's components."""
import time

import numpy as np

CFG5 = dict(update_travel_dist=0.05, searcher_candidates=64, searcher_node_dist=12.0, searcher_travel_dist=10.0)


def make_trip(n_scans, seed=50000):
    from my_lidar_graph_slam_v2_b200 import synth
    rng = np.random.default_rng(seed)
    world = synth.corridor_world(rng)
    return synth.corridor_trajectory(world, n_scans, 0.1, rng)


def summarize(c, wall_s):
    front = c["t_latest_map"] + c["t_match"] + c["t_append"]
    return {
        "scans": int(c["scans_processed"]), "wall_s": wall_s, "scans_per_s": c["scans_processed"] / wall_s,
        "front_end_ms_per_scan": 1e3 * front / max(c["scans_processed"], 1),
        "latest_map_ms_per_scan": 1e3 * c["t_latest_map"] / max(c["scans_processed"], 1),
        "match_plus_final_ms_per_scan": 1e3 * c["t_match"] / max(c["scans_processed"], 1),
        "append_scan_ms_per_scan": 1e3 * c["t_append"] / max(c["scans_processed"], 1),
        "backend_steps": int(c["backend_steps"]), "detect_calls": int(c["backend_steps_with_candidates"]),
        "loop_queries": int(c["loop_queries"]), "loops_detected": int(c["loops_detected"]),
        "detect_s": c["t_detect"],
        "detect_calls_per_s": c["backend_steps_with_candidates"] / c["t_detect"] if c["t_detect"] > 0 else None,
        "detect_queries_per_s": c["loop_queries"] / c["t_detect"] if c["t_detect"] > 0 else None,
        "optimizer_calls": int(c["optimizer_calls"]), "degenerations": int(c["degenerations"]),
    }


def run_gpu(trip, device=0, **over):
    from my_lidar_graph_slam_v2_b200 import hostapi, slam_settings
    ctx = hostapi.Context(device)
    slam = hostapi.SlamPipeline(ctx, slam_settings.pack(**dict(CFG5, **over)))
    t0 = time.perf_counter()
    slam.run(trip["angles"], trip["ranges"], trip["odom"], trip["stamps"], 0.01, 11.3, finish=True)
    wall = time.perf_counter() - t0
    out = summarize(slam.counters(), wall)
    nodes = slam.scan_nodes()
    out["local_maps"] = len(slam.local_maps())
    out["edges"] = len(slam.edges())
    true = trip["true"][:len(nodes)]
    # the estimate starts at the origin with heading 0; so does the true path relative to its first pose
    d = nodes[:, :2] - (true[:, :2] - true[0, :2])
    out["position_error_m"] = {"mean": float(np.hypot(d[:, 0], d[:, 1]).mean()), "end": float(np.hypot(*d[-1]))}
    slam.close()
    ctx.close()
    return out, nodes


def run_gpu_from_carmen(trip, log_path, device=0, metrics_path=None, **over):
    """The launcher's way in and out: the trajectory as a Carmen log (ROBOTLASER1 + ODOM records, written here),
    read back by the C++ CarmenLogReader, the loop run from its records (SlamPipeline::RunLog), the metrics saved
    as <metrics_path>.metric.json. Returns (summary, scan nodes, seconds spent reading the log)."""
    from my_lidar_graph_slam_v2_b200 import hostapi, slam_settings
    n_beams = trip["ranges"].shape[1]
    start, inc = -np.pi, 2.0 * np.pi / n_beams
    hostapi.write_carmen_log(log_path, trip["ranges"], trip["odom"], trip["stamps"], start, inc, 11.3)
    t0 = time.perf_counter()
    log = hostapi.CarmenLog(path=log_path)
    t_read = time.perf_counter() - t0
    ctx = hostapi.Context(device)
    slam = hostapi.SlamPipeline(ctx, slam_settings.pack(**dict(CFG5, **over)))
    if metrics_path:
        slam.record_metrics()
    t0 = time.perf_counter()
    slam.run_carmen(log, finish=True)
    wall = time.perf_counter() - t0
    out = summarize(slam.counters(), wall)
    out["log_read_s"] = t_read
    nodes = slam.scan_nodes()
    if metrics_path:
        slam.save_metrics(metrics_path)
    slam.close()
    log.close()
    ctx.close()
    return out, nodes
