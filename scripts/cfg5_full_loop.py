"""BASELINE configs[4]: the full SLAM loop on a synthetic 10k-scan closed trajectory (60 x 40 m corridor
world, drifting odometry). GPU arm: C++ SlamPipeline (every matcher and every map on the device, the
pose-graph optimiser behind its seam as the identity). Reference arm: the same loop on the compiled
reference's components (oracle/_ref), on a prefix of the trajectory.

    python scripts/cfg5_full_loop.py [n_scans] [n_reference_scans] [--carmen DIR]
prints one JSON object. With --carmen the GPU arm runs a second time from a Carmen log written into DIR and read back
by the C++ reader (the launcher's input format), saves DIR/cfg5.metric.json, and the two pose graphs are compared."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from my_lidar_graph_slam_v2_b200.full_loop import CFG5, make_trip, run_gpu, run_gpu_from_carmen, summarize


def run_reference(trip, n, kind="reference"):
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import slam_settings
    ref = pyoracle.load(kind)
    slam = ref.slam(slam_settings.pack(host_final_matchers=1, **CFG5))
    t0 = time.perf_counter()
    slam.run(trip["angles"], trip["ranges"][:n], trip["odom"][:n], trip["stamps"][:n], 0.01, 11.3, finish=True)
    wall = time.perf_counter() - t0
    out = summarize(slam.counters(), wall)
    nodes = slam.scan_nodes()
    slam.close()
    return out, nodes


def main():
    carmen_dir = None
    if "--carmen" in sys.argv:
        k = sys.argv.index("--carmen")
        carmen_dir = sys.argv[k + 1]
        del sys.argv[k:k + 2]
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    trip = make_trip(n)
    gpu, gnodes = run_gpu(trip)
    ref, rnodes = run_reference(trip, n_ref)
    m = min(len(gnodes), len(rnodes))
    out = {"workload": "cfg5: %d scans of 360 beams, 0.1 m apart, around a 60 x 40 m corridor loop; every scan matched; "
                       "local map every 2.5 m; loop detection every 2.5 m with up to 64 candidates" % n,
           "gpu": gpu, "reference": dict(ref, sample="the first %d scans of the same trajectory, 1 thread" % n_ref),
           "ratio": {"scans_per_s": gpu["scans_per_s"] / ref["scans_per_s"],
                     "detect_queries_per_s": (gpu["detect_queries_per_s"] / ref["detect_queries_per_s"])
                     if gpu["detect_queries_per_s"] and ref["detect_queries_per_s"] else None},
           "prefix_agreement": {"scans": m, "max_abs_pose_difference": float(np.abs(gnodes[:m, :3] - rnodes[:m, :3]).max())},
           "optimizer": "identity behind PoseGraphOptimizer (pose_graph.hpp); g2o is not in this image"}
    if carmen_dir:
        os.makedirs(carmen_dir, exist_ok=True)
        # the reader's angles are start + i * increment; the direct run above uses the generator's: compare the
        # two runs on the discrete structure and to 1e-9 on the poses
        cg, cnodes = run_gpu_from_carmen(trip, os.path.join(carmen_dir, "cfg5.log"),
                                         metrics_path=os.path.join(carmen_dir, "cfg5"))
        doc = json.load(open(os.path.join(carmen_dir, "cfg5.metric.json")))
        out["gpu_from_carmen_log"] = dict(
            cg, log_bytes=os.path.getsize(os.path.join(carmen_dir, "cfg5.log")),
            same_scan_nodes=len(cnodes) == len(gnodes),
            max_abs_pose_difference_vs_direct=float(np.abs(cnodes[:, :3] - gnodes[:len(cnodes), :3]).max()),
            metric_ids=len(doc["ValueSequences"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
