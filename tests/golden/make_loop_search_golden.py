"""Generates tests/golden/loop_search_vectors.json from the compiled reference (oracle/_ref):
LoopSearcherNearest::Search on seeded pose-graph summaries. Run in the build container."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from my_lidar_graph_slam_v2_b200 import synth          # noqa: E402
from oracle import pyoracle                              # noqa: E402

CASES = [(5000 + k, thr, node, cand) for k, (thr, node, cand) in enumerate([
    (10.0, 2.0, 2), (5.0, 3.0, 8), (20.0, 5.0, 64), (1.0, 1.0, 4), (30.0, 4.0, 256), (8.0, 2.5, 1),
    (1000.0, 5.0, 8), (5.0, 0.2, 16)])]

if __name__ == "__main__":
    orc = pyoracle.load("reference")
    out = []
    for seed, travel, node, cand in CASES:
        g = synth.make_pose_graph_summary(seed, loop=seed % 2 == 0)
        res = orc.loop_search(travel_dist_threshold=travel, node_dist_threshold=node,
                              num_of_candidate_nodes=cand, **g)
        out.append({"seed": seed, "travel": travel, "node": node, "cand": cand,
                    "loop": seed % 2 == 0, "candidates": res})
        print(seed, len(res))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "loop_search_vectors.json"), "w") as f:
        json.dump({"loop_search": out}, f, indent=1)
