"""Settings of the full SLAM loop (host/include/csm_host/slam_pipeline.hpp: SlamSettings), in the order
csm_host_slam_create reads them; defaults are the reference's launcher_settings_default.json."""
import numpy as np

FIELDS = [
    ("resolution", 0.05), ("patch_size", 16), ("scans_for_latest_map", 10), ("local_map_travel_dist", 2.5),
    ("overlapped_scans", 10), ("usable_range_min", 0.01), ("usable_range_max", 20.0), ("prob_hit", 0.62),
    ("prob_miss", 0.46),
    ("update_travel_dist", 0.5), ("update_angle", 0.5), ("update_time", 5.0), ("loop_detection_threshold", 2.5),
    ("degeneration_threshold", 10.0), ("odometry_covariance_scale", 1e2), ("fuse_odometry_covariance", 0),
    ("initial_x", 0.0), ("initial_y", 0.0), ("initial_theta", 0.0),
    ("rt_low_resolution", 5), ("rt_range_x", 0.25), ("rt_range_y", 0.25), ("rt_range_theta", 0.5),
    ("final_iterations", 10), ("final_convergence", 1e-4), ("final_lambda", 1e-4), ("covariance_scale", 1e4),
    ("searcher_travel_dist", 10.0), ("searcher_node_dist", 5.0), ("searcher_candidates", 2),
    ("bb_node_height_max", 6), ("bb_range_x", 2.5), ("bb_range_y", 2.5), ("bb_range_theta", 0.5),
    ("score_threshold", 0.55), ("known_rate_threshold", 0.6),
    ("host_final_matchers", 0),
]


def pack(**overrides):
    unknown = set(overrides) - {k for k, _ in FIELDS}
    assert not unknown, unknown
    return np.array([float(overrides.get(k, d)) for k, d in FIELDS], dtype=np.float64)
