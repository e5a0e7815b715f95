"""The drop-in on the reference's OWN types: tests/integration/csm_gpu_adapter.hpp holds classes derived
from the reference's ScanMatcher (mapping/scan_matcher.hpp:89-117) and LoopDetector
(mapping/loop_detector.hpp:97-116), compiled against the reference's headers (oracle/Makefile, target
`adapter`). The driver calls them and the reference's CPU classes through the base-class virtuals with
the reference's ScanMatchingQuery / LoopDetectionQueryVector and returns both outcomes.

Bar: what the reference returns -- found flag, estimated pose, normalized cost, covariance, the
LoopDetectionResult vector -- identical (the epilogue and the CPU final matcher ARE the reference's code,
evaluated at the pose the GPU found), 1e-5 relative where the final matcher runs on the device."""
import ctypes as C
import os

import numpy as np
import pytest

from my_lidar_graph_slam_v2_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def adapter():
    from oracle import pyoracle
    path = pyoracle._PATHS["adapter"]
    if not os.path.exists(path):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref/libcsm_adapter.so not built and /root/reference absent")
        pyoracle.build("adapter")
    lib = C.CDLL(path)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.adp_match_case.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                   dp, dp, C.c_int, dp, dp, C.c_int, dp, dp, dp, dp, ip]
    lib.adp_loop_case.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, dp, dp, C.POINTER(C.c_int32),
                                  dp, dp, dp, dp, C.c_int, C.c_int, dp, C.c_double, C.c_double, C.c_int, C.c_int,
                                  ip, C.POINTER(C.c_int32), dp, ip, C.POINTER(C.c_int32), dp, ip]
    return lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("kind,param,rng,step", [
    (0, 5, synth.CFG1["rng"], (0, 0, 0)),
    (1, 5, synth.CFG2["rng"], (0, 0, 0)),
    (1, 6, synth.CFG3["rng"], (0, 0, 0)),
    (2, 0, (0.4, 0.4, 0.06), (0.05, 0.05, 0.004)),
])
@pytest.mark.parametrize("seed", [5100, 5101])
def test_scan_matcher_virtuals(adapter, kind, param, rng, step, seed):
    """ScanMatcher::OptimizePose(const ScanMatchingQuery&) of the reference's CPU matcher and of the GPU
    class created by the factory string, on the same GridMap / ScanData objects."""
    case = synth.case_for(synth.CFG1, seed)
    s = case.submap
    grid = np.ascontiguousarray(s.grid)
    rel = (0.1, -0.03, 0.2) if seed % 2 else (0.0, 0.0, 0.0)
    a, ap = _d(case.angles)
    r, rp = _d(case.ranges)
    init, ip_ = _d(case.init_pose)
    relv, relp = _d(rel)
    rg, rgp = _d(rng)
    st, stp = _d(step)
    cpu = np.zeros(14)
    gpu = np.zeros(14)
    flags = C.c_int(0)
    assert adapter.adp_match_case(kind, grid.ctypes.data, grid.shape[0], grid.shape[1], s.res, s.off_x, s.off_y,
                                  ap, rp, len(a), ip_, relp, param, rgp, stp,
                                  cpu.ctypes.data_as(C.POINTER(C.c_double)), gpu.ctypes.data_as(C.POINTER(C.c_double)),
                                  C.byref(flags)) == 0
    assert flags.value == 0
    assert cpu[0] == gpu[0] == 1.0
    # the epilogue is the reference's own code at the pose the device found: bit-identical
    assert np.array_equal(cpu, gpu), (cpu, gpu)


@pytest.mark.parametrize("device_refiner", [0, 1])
def test_loop_detector_virtual(adapter, device_refiner):
    """LoopDetector::Detect(const LoopDetectionQueryVector&) of the reference's LoopDetectorBranchBound and
    of LoopDetectorBranchBoundGPU on the reference's own LocalMap / ScanNode / LocalMapNode objects: the
    same LoopDetectionResultVector (two calls each: first touch and cached)."""
    batch = synth.make_loop_batch(5200, n_maps=12, true_fraction=0.5, map_id_base=40)
    grids = np.ascontiguousarray(np.stack([m.grid for m in batch.submaps]))
    nq = len(batch.submaps)
    offx, oxp = _d([m.off_x for m in batch.submaps])
    offy, oyp = _d([m.off_y for m in batch.submaps])
    ids = np.ascontiguousarray(batch.map_ids, dtype=np.int32)
    mp, mpp = _d(batch.map_poses)
    sp, spp = _d(batch.scan_poses)
    a, ap = _d(batch.angles[0])
    r, rp = _d(batch.ranges[0])
    rg, rgp = _d(synth.CFG3["rng"])
    out = {}
    for who in ("cpu", "gpu"):
        out[who] = dict(n=C.c_int(0), ids=np.zeros(2 * nq, dtype=np.int32), vals=np.zeros(15 * nq))
    flags = C.c_int(0)
    assert adapter.adp_loop_case(
        nq, grids.ctypes.data, 512, 512, batch.submaps[0].res, oxp, oyp, ids.ctypes.data_as(C.POINTER(C.c_int32)),
        mpp, spp, ap, rp, len(a), 6, rgp, synth.CFG3["thr"][0], synth.CFG3["thr"][1], device_refiner, 2,
        C.byref(out["cpu"]["n"]), out["cpu"]["ids"].ctypes.data_as(C.POINTER(C.c_int32)),
        out["cpu"]["vals"].ctypes.data_as(C.POINTER(C.c_double)),
        C.byref(out["gpu"]["n"]), out["gpu"]["ids"].ctypes.data_as(C.POINTER(C.c_int32)),
        out["gpu"]["vals"].ctypes.data_as(C.POINTER(C.c_double)), C.byref(flags)) == 0
    n = out["cpu"]["n"].value
    assert n == out["gpu"]["n"].value >= 3
    assert flags.value == 0
    assert np.array_equal(out["cpu"]["ids"][:2 * n], out["gpu"]["ids"][:2 * n])
    cv, gv = out["cpu"]["vals"][:15 * n].reshape(n, 15), out["gpu"]["vals"][:15 * n].reshape(n, 15)
    if device_refiner:
        assert np.allclose(cv[:, :6], gv[:, :6], rtol=1e-5, atol=0.0)        # refined poses: north_star tolerance
        assert np.allclose(cv[:, 6:], gv[:, 6:], rtol=1e-4, atol=1e-12)
    else:
        assert np.array_equal(cv, gv)        # the reference's own final matcher from the same coarse pose
