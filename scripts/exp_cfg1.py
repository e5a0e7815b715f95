"""Experiment: latency breakdown of one real-time correlative match (cfg1) through the C ABI."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from my_lidar_graph_slam_v2_b200 import capi, matchers, synth, hostapi

h = capi.Handle(0)
case = synth.case_for(synth.CFG1, 41000)
s = case.submap
scan = matchers.ScanData(case.angles, case.ranges)
sensor = matchers.compound(tuple(case.init_pose), scan.relative_sensor_pose)
step = matchers.compute_search_step(s.res, scan.ranges)
win = matchers.search_window(synth.CFG1["rng"], step)
grid = np.ascontiguousarray(s.grid)
blocks, index, br, bc = synth.dense_to_blocks(grid)
print("blocks", len(index), "of", br * bc)
REP = int(sys.argv[1]) if len(sys.argv) > 1 else 200

def phase(fn, sync=True):
    t0 = time.perf_counter()
    for _ in range(REP):
        fn()
        if sync: h.synchronize()
    return (time.perf_counter() - t0) / REP * 1e6

up = lambda: h.upload_grid(5, grid, s.res, s.off_x, s.off_y)
upb = lambda: h.upload_grid_blocks(5, blocks, index, 4, br, bc, s.res, s.off_x, s.off_y)
def coarse():
    h.drop_pyramids([5]); h.build_coarse(5, 5)
mt = lambda: h.match_rt(5, scan.angles, scan.ranges, sensor, 5, win, step, (0.0, 0.0))
up(); coarse(); mt(); h.synchronize()
print("upload dense+sync %.1f us | upload blocks+sync %.1f | coarse+sync %.1f | match_rt %.1f" % (
    phase(up), phase(upb), phase(coarse), phase(mt, False)))
def full():
    up(); h.build_coarse(5, 5); return mt()
def fullb():
    upb(); h.build_coarse(5, 5); return mt()
print("full dense %.1f us | full blocks %.1f us" % (phase(full, False), phase(fullb, False)))
ctx = hostapi.Context(0)
f = lambda: ctx.match("rt", grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges, case.init_pose, 5, synth.CFG1["rng"])
f()
print("C++ adapter (incl. CPU cost/covariance) %.1f us" % phase(f, False))
t0 = time.perf_counter()
for _ in range(REP):
    hostapi.cost(grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges, case.true_pose)
print("CPU epilogue alone %.1f us" % ((time.perf_counter() - t0) / REP * 1e6))
bbm = lambda: h.match_bb(5, scan.angles, scan.ranges, sensor, 5, matchers.search_window(synth.CFG2["rng"], step), step, (0.0, 0.0))
def pyr():
    h.drop_pyramids([5]); h.build_pyramid(5, 5)
pyr(); bbm()
print("cfg2: pyramid+sync %.1f us | match_bb %.1f us" % (phase(pyr), phase(bbm, False)))
# the bench's single_scan legs (block-sparse map through the C++ plugin)
def rtb():
    return ctx.match_blocks("rt", blocks.copy(), index, 4, s.grid.shape, s.res, (s.off_x, s.off_y), case.angles,
                            case.ranges, case.init_pose, 5, synth.CFG1["rng"])
def rtb_nocopy():
    return ctx.match_blocks("rt", blocks, index, 4, s.grid.shape, s.res, (s.off_x, s.off_y), case.angles,
                            case.ranges, case.init_pose, 5, synth.CFG1["rng"])
def bbb():
    return ctx.match_blocks("bb", blocks.copy(), index, 4, s.grid.shape, s.res, (s.off_x, s.off_y), case.angles,
                            case.ranges, case.init_pose, 5, synth.CFG2["rng"])
for name, fn in (("rt blocks+copy", rtb), ("rt blocks", rtb_nocopy), ("bb blocks+copy", bbb), ("rt blocks+copy", rtb)):
    fn()
    for rep in range(3):
        print("plugin %s: %.1f us" % (name, phase(fn, False)))
hh = capi.Handle.from_pointer(ctx.handle(), 0)
hh.set_option("timing", 2)
t0 = time.perf_counter(); rtb_nocopy(); t1 = time.perf_counter()
print("one rt call wall %.1f us" % ((t1 - t0) * 1e6))
for name, ms in hh.timings():
    print("  %8.3f ms  %s" % (ms, name))
hh.set_option("timing", 0)
ctx.set_device_epilogue(True)
for name, fn in (("rt blocks, device epilogue", rtb_nocopy), ("bb blocks+copy, device epilogue", bbb)):
    fn()
    for rep in range(2):
        print("plugin %s: %.1f us" % (name, phase(fn, False)))
