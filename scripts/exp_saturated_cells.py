"""How the reference's matchers treat cells at 65535 (it reads past its 65535-entry tables there) on maps
the builder makes, against the device with (A) 65535 kept (p = 0.999) and (B) 65535 taken as unknown."""
import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import pyoracle
from my_lidar_graph_slam_v2_b200 import hostapi, synth
ref = pyoracle.load("reference")
ctx = hostapi.Context(0)
stats = {}
for seed in range(30):
    rng = np.random.default_rng(9000 + seed)
    room = synth.make_room(rng)
    p0 = synth.random_pose_in_room(room, rng)
    ob = ref.map_builder()
    traj = []
    for k in range(11):
        p = p0 + np.array([0.06, 0.025, 0.012]) * k
        a, r = synth.raycast(room, p, 360, 0.01, 11.4, rng)
        traj.append((p, a, r))
    for p, a, r in traj[:10]:
        ob.append(p, a, r)
    dense, _, off, _, _ = ob.latest()
    nsat = int((dense == 65535).sum())
    g = ref.grid(dense, 0.05, off[0], off[1])
    canon = dense.copy(); canon[canon == 65535] = 0
    p, a, r = traj[10]
    init = p + np.array([0.07, -0.05, 0.02])
    for kind, win, par in (("rt", synth.CFG1["rng"], 5), ("bb", synth.CFG2["rng"], 5)):
        exp = (ref.match_rt if kind == "rt" else ref.match_bb)(g, a, r, init, par, win)
        for name, m in (("kept", dense), ("unknown", canon)):
            got = ctx.match(kind, m, 0.05, off, a, r, init, par, win)
            same = (got.best_x, got.best_y, got.best_t) == (exp.best_x, exp.best_y, exp.best_t)
            exact = same and got.score == exp.score
            s = stats.setdefault((kind, name), [0, 0, 0])
            s[0] += 1; s[1] += same; s[2] += exact
    if seed < 3:
        print("seed", seed, "cells at 65535:", nsat, "of", int((dense > 0).sum()), "known")
for k, v in sorted(stats.items()):
    print(k, "cases %d, same best index %d, same index and score bits %d" % tuple(v))
