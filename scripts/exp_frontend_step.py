"""Front-end step on the resident latest map: UpdateLatestMap + RT match + final matcher, timed apart."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
from my_lidar_graph_slam_v2_b200 import hostapi, synth
REFINE = (10, 1e-4, 1e-4)
rng_t = np.random.default_rng(41001)
room = synth.make_room(rng_t)
p0 = synth.random_pose_in_room(room, rng_t)
traj = []
for k in range(40):
    p = p0 + np.array([0.06, 0.025, 0.012]) * k
    a, r = synth.raycast(room, p, 360, 0.01, 11.4, rng_t)
    traj.append((p, a, r))
ctx = hostapi.Context(0)
ctx.set_device_final_matcher(*REFINE)
mb = hostapi.MapBuilder(ctx)
for p, a, r in traj[:10]:
    mb.append(p, a, r)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
for phase in ("warm", "timed"):
    tu = tm = 0.0
    k = 10
    for i in range(reps):
        k = 10 + (k - 9) % 30
        p, a, r = traj[k]
        t0 = time.perf_counter()
        mb.append(p, a, r)
        t1 = time.perf_counter()
        mb.match_rt(a, r, p + np.array([0.07, -0.05, 0.02]), 5, synth.CFG1["rng"])
        t2 = time.perf_counter()
        tu += t1 - t0; tm += t2 - t1
print("update %.1f us, match+final %.1f us per scan" % (tu / reps * 1e6, tm / reps * 1e6))
