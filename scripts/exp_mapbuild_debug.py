import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from oracle import pyoracle
from my_lidar_graph_slam_v2_b200 import hostapi, synth
from test_gpu_mapbuild import _trajectory
ref = pyoracle.load("reference")
for n_beams, rel in [(1080, (0.12, -0.04, 0.3)), (1080, (0, 0, 0)), (720, (0.12, -0.04, 0.3))]:
    _, traj = _trajectory(7002, 14, n_beams)
    ctx = hostapi.Context(0)
    mb = hostapi.MapBuilder(ctx); ob = ref.map_builder()
    for k, (p, a, r) in enumerate(traj):
        mb.append(p, a, r, rel); ob.append(p, a, r, rel)
        d, o = mb.latest(), ob.latest()
        bad = np.argwhere(d[0] != o[0])
        if len(bad):
            print(n_beams, rel, "scan", k, "bad", len(bad), d[0].shape, d[2], d[3])
            for b in bad[:4]:
                r0, c0 = b
                print(" cell", b, "dev", d[0][r0, c0], "ref", o[0][r0, c0])
                print(" dev nb\n", d[0][r0-2:r0+3, c0-2:c0+3], "\n ref nb\n", o[0][r0-2:r0+3, c0-2:c0+3])
            # sensor cells
            mp = d[3]
            for kk in range(max(0, k - 9), k + 1):
                pp = traj[kk][0]
                c, s = np.cos(pp[2]), np.sin(pp[2])
                gx = pp[0] + c * rel[0] - s * rel[1]; gy = pp[1] + s * rel[0] + c * rel[1]
                dx, dy = gx - mp[0], gy - mp[1]
                cm, sm = np.cos(mp[2]), np.sin(mp[2])
                lx = cm * dx + sm * dy; ly = -sm * dx + cm * dy
                print("  scan", kk, "sensor cell row", int(np.floor((ly - d[2][1]) / 0.05)), "col", int(np.floor((lx - d[2][0]) / 0.05)))
            break
    mb.close(); ctx.close()
