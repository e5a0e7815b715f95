"""Shared helpers of the parity tests."""
import hashlib
import json
import os

import numpy as np

from my_lidar_graph_slam_v2_b200 import matchers, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCORE_RTOL = 1e-5          # north_star: float scores within 1e-5 relative


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def grid_of(case_or_submap):
    s = getattr(case_or_submap, "submap", case_or_submap)
    return matchers.GridMap(s.grid, s.res, (s.off_x, s.off_y))


def assert_match(dev, orc, what, exact_score=True, flags_ok=0):
    """dev: capi.CsmResult, orc: OrcResult or golden dict. Every compared result must come without
    flags (FP guard band, key tie, inadmissible edge) unless the test allows some in `flags_ok`."""
    o = orc if isinstance(orc, dict) else orc.asdict()
    assert dev.flags & ~flags_ok == 0, "%s: flags %d" % (what, dev.flags)
    assert dev.found == o["found"], "%s: found %d vs %d" % (what, dev.found, o["found"])
    if o["found"] or o.get("compare_unfound", True):
        got = (dev.best_x, dev.best_y, dev.best_t)
        exp = (o["best_x"], o["best_y"], o["best_t"])
        assert got == exp, "%s: best index %s vs %s" % (what, got, exp)
    if o["found"]:
        assert dev.sum_value == o["sum_value"], "%s: sum_value" % what
        assert dev.n_known == o["n_known"], "%s: n_known" % what
        exp_score = float.fromhex(o["score"]) if isinstance(o["score"], str) else o["score"]
        assert abs(dev.normalized_score - exp_score) <= SCORE_RTOL * abs(exp_score), "%s: score" % what
        if exact_score:
            assert dev.normalized_score == exp_score, "%s: score not bit-identical" % what


def as_matchers_read(grid):
    """The matchers' view of a map: a cell at 65535 reads as unknown, like in the compiled reference, whose
    65535-entry value tables end one short of it (grid_values.cpp:32-35; csrc/csm_kernels.cuh,
    k_saturated_unknown)."""
    g = np.array(grid, copy=True)
    g[g == 65535] = 0
    return g
