"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The reference ships no tests or known-answer vectors, so
these fixtures are outputs of the reference's own code on seeded synthetic
inputs; inputs are regenerated from the seeds by synth.py and guarded by the
stored input digests.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle  # noqa: E402
from my_lidar_graph_slam_v2_b200 import synth  # noqa: E402
from helpers import sha  # noqa: E402


def res_dict(r):
    d = r.asdict()
    d["score"] = float(d["score"]).hex()
    return d


def small_cfg(cfg, rows=256, cols=256):
    c = dict(cfg)
    c["rows"], c["cols"] = rows, cols
    return c


def main():
    ref = pyoracle.load("reference")
    out = {"matches": [], "pyramids": [], "loop": []}

    # --- single-scan matchers ------------------------------------------------
    for seed in range(6):
        case = synth.case_for(synth.CFG1, 1000 + seed)
        s = case.submap
        g = ref.grid(s.grid, s.res, s.off_x, s.off_y)
        base = dict(seed=1000 + seed, grid_sha=sha(s.grid), scan_sha=sha(case.ranges))
        for thr in ((0.0, 0.0), (0.4, 0.5)):
            r = ref.match_rt(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"], thr)
            out["matches"].append(dict(base, kind="rt", cfg="CFG1", low_res=5, thr=thr, expect=res_dict(r)))
        r = ref.match_rt(g, case.angles, case.ranges, case.init_pose, 3, (0.6, 0.4, 0.1), (0.0, 0.0))
        out["matches"].append(dict(base, kind="rt", cfg="CFG1", low_res=3, rng=(0.6, 0.4, 0.1),
                                   thr=(0.0, 0.0), expect=res_dict(r)))
        r = ref.match_bb(g, case.angles, case.ranges, case.init_pose, 5, synth.CFG2["rng"], (0.0, 0.0))
        out["matches"].append(dict(base, kind="bb", cfg="CFG1", hmax=5, rng=synth.CFG2["rng"],
                                   thr=(0.0, 0.0), expect=res_dict(r)))
        r = ref.match_bb(g, case.angles, case.ranges, case.init_pose, 6, synth.CFG3["rng"], synth.CFG3["thr"])
        out["matches"].append(dict(base, kind="bb", cfg="CFG1", hmax=6, rng=synth.CFG3["rng"],
                                   thr=synth.CFG3["thr"], expect=res_dict(r)))
        r = ref.match_grid(g, case.angles, case.ranges, case.init_pose, (0.4, 0.4, 0.06),
                           (0.05, 0.05, 0.004), (0.0, 0.0))
        out["matches"].append(dict(base, kind="grid", cfg="CFG1", rng=(0.4, 0.4, 0.06),
                                   step=(0.05, 0.05, 0.004), thr=(0.0, 0.0), expect=res_dict(r)))
        r = ref.match_grid(g, case.angles, case.ranges, case.init_pose, (0.3, 0.3, 0.04),
                           (0.03, 0.07, 0.005), (0.3, 0.5))
        out["matches"].append(dict(base, kind="grid", cfg="CFG1", rng=(0.3, 0.3, 0.04),
                                   step=(0.03, 0.07, 0.005), thr=(0.3, 0.5), expect=res_dict(r)))

    # --- pyramids -----------------------------------------------------------
    for seed, (rows, cols) in enumerate([(512, 512), (256, 384), (64, 48), (32, 32), (16, 160)]):
        rng = np.random.default_rng(7000 + seed)
        if rows >= 256:
            sub = synth.rasterize(synth.make_room(rng, 8.0, 6.0, 1.0), rng, rows, cols, 0.05)
            grid = sub.grid
        else:
            grid = rng.integers(0, 65535, size=(rows, cols), dtype=np.uint16)
            grid[rng.random((rows, cols)) < 0.5] = 0
        g = ref.grid(grid, 0.05, -1.0, -2.0)
        pyr = g.pyramid(6)
        coarse = {str(w): sha(g.precompute(w)) for w in (1, 2, 3, 5, 7, 8)}
        out["pyramids"].append(dict(seed=7000 + seed, rows=rows, cols=cols, grid_sha=sha(grid),
                                    levels=[sha(pyr[h]) for h in range(7)], coarse=coarse))

    # --- loop detection batch --------------------------------------------------
    batch = synth.make_loop_batch(3000, n_maps=24, true_fraction=0.4)
    grids = [ref.grid(s.grid, s.res, s.off_x, s.off_y) for s in batch.submaps]
    det = ref.loop_detector(6, synth.CFG3["rng"], synth.CFG3["thr"], 1)
    res, _ = det.detect(grids, batch.map_ids, batch.map_poses, batch.scan_idx, batch.scan_poses,
                        batch.angles, batch.ranges)
    out["loop"].append(dict(seed=3000, n_maps=24, true_fraction=0.4, hmax=6,
                            grid_sha=sha(np.stack([s.grid for s in batch.submaps])),
                            expect=[res_dict(r) for r in res]))

    # --- linear-solver refiner (the loop detector's default final matcher) ---------------
    # outputs of the reference's scan_matcher_linear_solver.cpp compiled against the Eigen shim
    # (column-pivoting Householder QR for the 3x3 solve, oracle/ref_shim/Eigen/Core)
    refine = []
    for seed in range(8):
        case = synth.case_for(synth.CFG1, 1500 + seed)
        s = case.submap
        g = ref.grid(s.grid, s.res, s.off_x, s.off_y)
        rng = np.random.default_rng(1500 + seed)
        init = case.true_pose + rng.uniform(-1.0, 1.0, size=3) * np.array([0.05, 0.05, 0.02])
        rel = (0.1, -0.03, 0.2) if seed % 2 else (0.0, 0.0, 0.0)
        r = ref.refine(g, case.angles, case.ranges, init, rel, 10, 1e-4, 1e-4)
        refine.append(dict(seed=1500 + seed, grid_sha=sha(s.grid), scan_sha=sha(case.ranges),
                           init=[float(v).hex() for v in init], rel=list(rel), iterations=r.n_processed,
                           est_pose=[float(v).hex() for v in r.est_pose], norm_cost=float(r.norm_cost).hex(),
                           cov=[float(v).hex() for v in r.cov]))
    with open(os.path.join(HERE, "refine_vectors.json"), "w") as f:
        json.dump({"refine": refine}, f, indent=1)

    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["matches"]), "match vectors,", len(out["pyramids"]), "pyramids,",
          len(out["loop"]), "loop batches")


if __name__ == "__main__":
    main()
