"""Full-size BASELINE configs[3] golden vector from the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_cfg4_golden.py [seed ...]

Runs the reference's ScanMatcherGridSearch::OptimizePose (scan_matcher_grid_search.cpp:84-178) once
per seed on the full cfg4 case: 1080 beams, 1280 x 1280 map at 0.025 m, window 4 m x 4 m x 60 deg at
0.025 m / 0.1 deg = 161 x 161 x 600 candidates. One run takes ~9 minutes on one core, so it never
runs in CI: the winner (index, sum, known count, hex score) is committed as
tests/golden/cfg4_vectors.json and test_full_size_cfg4_grid_search asserts equality with it.
Seeds run in parallel processes (the reference is single-threaded).
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def one(seed):
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import synth
    from helpers import sha
    ref = pyoracle.load("reference")
    case = synth.case_for(synth.CFG4, seed)
    s = case.submap
    g = ref.grid(s.grid, s.res, s.off_x, s.off_y)
    t0 = time.time()
    r = ref.match_grid(g, case.angles, case.ranges, case.init_pose, synth.CFG4["rng"], synth.CFG4["step"],
                       synth.CFG4["thr"])
    d = r.asdict()
    d["score"] = float(d["score"]).hex()
    d["known_rate"] = float(d["known_rate"]).hex()
    for k in ("best_sensor_pose", "est_pose", "cov"):
        d[k] = [float(v).hex() for v in d[k]]
    d["norm_cost"] = float(d["norm_cost"]).hex()
    return dict(seed=seed, cfg="CFG4", grid_sha=sha(s.grid), scan_sha=sha(case.ranges),
                init_pose=[float(v).hex() for v in case.init_pose], reference_seconds=round(time.time() - t0, 1),
                expect=d)


def main():
    seeds = [int(a) for a in sys.argv[1:]] or [44001, 44002]
    with ProcessPoolExecutor(max_workers=len(seeds)) as ex:
        out = list(ex.map(one, seeds))
    with open(os.path.join(HERE, "cfg4_vectors.json"), "w") as f:
        json.dump({"cfg4": out}, f, indent=1)
    print("wrote", len(out), "cfg4 vectors:", [(o["seed"], o["reference_seconds"]) for o in out])


if __name__ == "__main__":
    main()
