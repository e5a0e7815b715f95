"""Generates tests/golden/hill_climb_vectors.json from the compiled reference (oracle/_ref):
ScanMatcherHillClimbing::OptimizePose over CostSquareError on seeded cases. Run in the build container."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from helpers import sha                                   # noqa: E402
from my_lidar_graph_slam_v2_b200 import synth            # noqa: E402
from oracle import pyoracle                               # noqa: E402

PARAMS = [(0.1, 0.1, 100, 5), (0.05, 0.02, 30, 3), (0.2, 0.05, 8, 10), (0.01, 0.01, 100, 0)]
# CostGreedyEndpoint: MapResolution, HitAndMissedDist, OccupancyThreshold, KernelSize, ScalingFactor,
# StandardDeviation (launcher_settings_default.json:2-9, and a wider kernel)
GREEDY = [(0.05, 0.075, 0.1, 1, 1.0, 0.05), (0.05, 0.1, 0.3, 2, 2.0, 0.08)]

if __name__ == "__main__":
    ref = pyoracle.load("reference")
    out = []
    for k in range(8):
        seed = 1600 + k
        case = synth.case_for(synth.CFG1, seed)
        s = case.submap
        g = ref.grid(s.grid, s.res, s.off_x, s.off_y)
        rng = np.random.default_rng(seed)
        init = case.true_pose + rng.uniform(-1.0, 1.0, size=3) * np.array([0.08, 0.08, 0.03])
        rel = (0.1, -0.03, 0.2) if k % 2 else (0.0, 0.0, 0.0)
        lin, ang, iters, refs = PARAMS[k % len(PARAMS)]
        greedy = GREEDY[(k // 2) % 2] if k >= 4 else None
        r = ref.hill_climb(g, case.angles, case.ranges, init, rel, lin, ang, iters, refs, greedy)
        out.append(dict(seed=seed, grid_sha=sha(s.grid), init=[float(v).hex() for v in init], rel=list(rel),
                        params=[lin, ang, iters, refs], greedy=list(greedy) if greedy else None,
                        iterations=r.n_processed, refinements=r.n_ignored,
                        est_pose=[float(v).hex() for v in r.est_pose], norm_cost=float(r.norm_cost).hex(),
                        cov=[float(v).hex() for v in r.cov]))
        print(seed, r.n_processed, r.n_ignored, list(r.est_pose))
    with open(os.path.join(HERE, "hill_climb_vectors.json"), "w") as f:
        json.dump({"hill_climb": out}, f, indent=1)
