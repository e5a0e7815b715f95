"""CPU suite: host-side logic of the plugin mirror and the C-ABI surface (no compute calls)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from my_lidar_graph_slam_v2_b200 import build, capi, matchers, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """libcsm_b200.so loads without a GPU and exports every function include/csm_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "csm_b200.h")).read()
    declared = set(re.findall(r"\b(csm_[a-z_0-9]+)\s*\(", header))
    declared -= {"csm_context"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    lib = capi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.csm_version() >= 100


def test_no_cpu_fallback_without_device():
    """Without a usable CUDA device csm_create fails loudly (there is no CPU path)."""
    lib = capi.load()
    if lib.csm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.CsmError):
        capi.Handle(0)


def test_struct_layouts_match_header(tmp_path):
    """The header compiles as plain C (no C++, no torch types across the boundary) and the ctypes
    mirrors in capi.py have the sizes and field offsets the C compiler gives the structs."""
    import subprocess
    assert C.sizeof(capi.CsmResult) == 48
    assert C.sizeof(capi.CsmLoopQuery) == 96
    assert capi.CsmResult.sum_value.offset == 16 and capi.CsmResult.normalized_score.offset == 32
    probes = [("csm_result", capi.CsmResult, ["found", "sum_value", "normalized_score", "n_ignored"]),
              ("csm_loop_query", capi.CsmLoopQuery, ["map_id", "sensor_pose", "win_x", "step_x", "known_thr"]),
              ("csm_refine_params", capi.CsmRefineParams, ["max_iterations", "convergence_threshold",
                                                           ("lambda", "lambda_"), "covariance_scale"]),
              ("csm_refined", capi.CsmRefined, ["pose", "covariance", "initial_cost", "final_cost",
                                                ("lambda", "lambda_"), "iterations", "valid"]),
              ("csm_refine_query", capi.CsmRefineQuery, ["map_id", "scan_id", "sensor_pose"])]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "csm_b200.h"', "int main(void) {"]
    for name, _, fields in probes:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f in fields:
            cf = f[0] if isinstance(f, tuple) else f
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, cf, name, cf))
    lines += ["return 0; }"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True,
                                                 check=True).stdout.splitlines())
    for name, cls, fields in probes:
        assert int(got[name]) == C.sizeof(cls), name
        for f in fields:
            cf, pf = f if isinstance(f, tuple) else (f, f)
            assert int(got["%s.%s" % (name, cf)]) == getattr(cls, pf).offset, (name, cf)


def test_product_does_not_touch_the_oracle():
    """Nothing in the product package or include/ references oracle/."""
    pkg = os.path.join(ROOT, "my_lidar_graph_slam_v2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), os.path.join(dirpath, f)


def test_search_step_and_window_expressions():
    """scan_matcher_correlative.cpp:141-146,255-274 on the synthetic cfg sizes."""
    ranges = np.array([3.0, 11.40, 7.5])
    step = matchers.compute_search_step(0.05, ranges)
    assert step[0] == step[1] == 0.05
    assert step[2] == math.acos(1.0 - 0.5 * (0.05 / 11.40) * (0.05 / 11.40))
    assert matchers.search_window(synth.CFG1["rng"], step) == (5, 5, 20)
    assert matchers.search_window(synth.CFG2["rng"], step) == (20, 20, 60)
    assert matchers.search_window(synth.CFG3["rng"], step) == (25, 25, 57)


def test_grid_search_offsets_accumulate_like_the_reference():
    d = matchers.grid_search_offsets(2.0, 0.025)
    assert len(d) == 161 and d[0] == -2.0
    acc = -2.0
    for v in d:
        assert v == acc
        acc += 0.025
    # accumulated rounding decides whether the last value is still <= r: whatever the
    # C loop of scan_matcher_grid_search.cpp:118-120 does, the helper does too
    dt = matchers.grid_search_offsets(30.0 * synth.DEG, 0.1 * synth.DEG)
    assert len(dt) in (600, 601) and dt[-1] <= 30.0 * synth.DEG < dt[-1] + 0.1 * synth.DEG


def test_pose_algebra_roundtrip():
    a, b = (1.5, -2.0, 0.7), (0.3, 0.1, -0.2)
    c = matchers.compound(a, b)
    back = matchers.inverse_compound(a, c)
    assert np.allclose(back, b, atol=1e-12)
    assert np.allclose(matchers.move_backward(c, b), a, atol=1e-12)


def test_synthetic_generators_are_deterministic_and_capped():
    a, b = synth.case_for(synth.CFG1, 77), synth.case_for(synth.CFG1, 77)
    assert np.array_equal(a.submap.grid, b.submap.grid) and np.array_equal(a.ranges, b.ranges)
    assert a.submap.grid.max() <= 65534          # reference LUT is one entry short
    assert a.submap.grid.shape == (512, 512) and a.ranges.max() == 11.40
    # all-unknown margin of at least 2^hmax cells on the low-index sides: a coarse cell that
    # starts at a negative index then covers only unknown cells, so the B&B bound stays
    # admissible there
    for seed in range(70, 90):
        g = synth.case_for(synth.CFG1, seed).submap.grid
        assert not g[:64].any() and not g[:, :64].any()
    lb = synth.make_loop_batch(5, n_maps=8)
    assert len(lb.submaps) == 8 and lb.map_poses.shape == (8, 3)


def test_build_flags_target_sm100a_only():
    flags = " ".join(build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags


def test_cpp_epilogue_cost_matches_reference_vectors():
    """The C++ adapter's CPU epilogue (cost / covariance at the winning pose) against the golden
    vectors of the reference: cost bit-identical, covariance within 1e-9 relative."""
    from helpers import load_golden
    from my_lidar_graph_slam_v2_b200 import hostapi
    gold = load_golden("reference_vectors.json")["matches"]
    checked = 0
    for m in gold:
        if not m["expect"]["found"] or checked >= 12:
            continue
        case = synth.case_for(synth.CFG1, m["seed"])
        s = case.submap
        e = m["expect"]
        nc, cov = hostapi.cost(s.grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges,
                               e["best_sensor_pose"])
        assert nc == e["norm_cost"], (m["kind"], m["seed"])
        assert np.allclose(cov, e["cov"], rtol=1e-9, atol=0.0)
        checked += 1
    assert checked == 12


def test_cpp_linear_solver_matches_reference_vectors():
    """The C++ host ScanMatcherLinearSolver (CPU) against vectors produced by the reference's
    scan_matcher_linear_solver.cpp: iteration count equal, refined pose and cost bit-identical
    (north_star asks for 1e-5 relative), covariance within 1e-9."""
    from helpers import load_golden, sha
    from my_lidar_graph_slam_v2_b200 import hostapi
    for e in load_golden("refine_vectors.json")["refine"]:
        case = synth.case_for(synth.CFG1, e["seed"])
        s = case.submap
        assert sha(s.grid) == e["grid_sha"]
        init = [float.fromhex(v) for v in e["init"]]
        out, lam = hostapi.refine(s.grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges, init,
                                  tuple(e["rel"]))
        exp_pose = [float.fromhex(v) for v in e["est_pose"]]
        assert out.best_t == e["iterations"]
        assert np.allclose(list(out.est_pose), exp_pose, rtol=1e-5, atol=0.0)      # the stated tolerance
        assert list(out.est_pose) == exp_pose                                      # what is actually achieved
        assert out.norm_cost == float.fromhex(e["norm_cost"])
        assert np.allclose(list(out.cov), [float.fromhex(v) for v in e["cov"]], rtol=1e-9, atol=0.0)
        assert 1e-8 <= lam <= 1e-4


def test_cpp_loop_searcher_matches_reference_vectors():
    """LoopSearcherNearest of the C++ mirror (host/src/loop_searcher.cpp, the caller that produces the
    loop-detection query batch) against vectors produced by the reference's loop_searcher_nearest.cpp:
    the same candidates in the same order (std::nth_element on the same sequence)."""
    from helpers import load_golden
    from my_lidar_graph_slam_v2_b200 import hostapi
    n_nonempty = 0
    for e in load_golden("loop_search_vectors.json")["loop_search"]:
        g = synth.make_pose_graph_summary(e["seed"], loop=e["loop"])
        got, dist = hostapi.loop_search(travel_dist_threshold=e["travel"], node_dist_threshold=e["node"],
                                        num_of_candidate_nodes=e["cand"], **g)
        assert got == [tuple(c) for c in e["candidates"]], e["seed"]
        assert len(got) <= e["cand"] and all(d < e["node"] ** 2 for d in dist)
        n_nonempty += bool(got)
    assert n_nonempty >= 4


def test_cpp_hill_climbing_matches_reference_vectors():
    """The C++ host ScanMatcherHillClimbing (CPU) against vectors produced by the reference's
    scan_matcher_hill_climbing.cpp over CostSquareError and over CostGreedyEndpoint: iteration and
    step-halving counts equal, estimated pose and cost bit-identical, covariance within 1e-9."""
    from helpers import load_golden, sha
    from my_lidar_graph_slam_v2_b200 import hostapi
    for e in load_golden("hill_climb_vectors.json")["hill_climb"]:
        case = synth.case_for(synth.CFG1, e["seed"])
        s = case.submap
        assert sha(s.grid) == e["grid_sha"]
        init = [float.fromhex(v) for v in e["init"]]
        lin, ang, iters, refs = e["params"]
        out = hostapi.hill_climb(s.grid, s.res, (s.off_x, s.off_y), case.angles, case.ranges, init,
                                 tuple(e["rel"]), lin, ang, int(iters), int(refs), greedy=e["greedy"])
        assert (out.best_t, out.best_x) == (e["iterations"], e["refinements"]), e["seed"]
        assert list(out.est_pose) == [float.fromhex(v) for v in e["est_pose"]], e["seed"]
        assert out.norm_cost == float.fromhex(e["norm_cost"])
        assert np.allclose(list(out.cov), [float.fromhex(v) for v in e["cov"]], rtol=1e-9, atol=0.0)


def test_adapter_on_reference_types_builds_and_exports():
    """The drop-in classes derived from the reference's ScanMatcher / LoopDetector
    (tests/integration/csm_gpu_adapter.hpp) compile against the reference's own headers and link with the
    C ABI library; the driver's entry points are there (no compute calls without a GPU)."""
    import ctypes as C
    from oracle import pyoracle
    path = pyoracle._PATHS["adapter"]
    if not os.path.exists(path) and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libcsm_adapter.so not built and /root/reference absent")
    if not os.path.exists(path):
        pyoracle.build("adapter")
    lib = C.CDLL(path)
    for sym in ("adp_match_case", "adp_loop_case"):
        assert hasattr(lib, sym)


def test_map_update_tables_equal_the_reference_cell_update():
    """The tables the device applies per map update (GridMapBuilderGPU::UpdateTable) against the reference's
    GridBinaryBayes::UpdateOddsUnchecked (grid_binary_bayes.cpp:302-321) for EVERY cell value, hit and miss odds of
    the launcher defaults and two others -- value 65535 included, where the compiled reference reads past its
    65535-entry odds table and drops the cell to ValueMin."""
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    if not (pyoracle.available("reference") or os.path.isdir("/root/reference")):
        pytest.skip("needs the compiled reference")
    ref = pyoracle.load("reference")
    for prob in (0.62, 0.46, 0.7, 0.4, 0.9):
        odds = prob / (1.0 - prob)
        exp = ref.update_table(odds)
        got = hostapi.map_update_table(odds)
        assert np.array_equal(got, exp), "p = %.2f: %d entries differ" % (prob, int((got != exp).sum()))
        assert got[65535] == 1
        cont = hostapi.map_update_table(odds, reference_table_end=False)
        assert np.array_equal(cont[:65535], exp[:65535]) and (cont[65535] == 65535) == (prob > 0.5)


def test_reference_reads_a_saturated_cell_as_unknown():
    """What the matchers' copy of a map reproduces (k_saturated_unknown): the compiled reference's
    ValueToProbability(65535) is one element past its 65535-entry table. It reads 0.0 (the unknown probability)
    when glibc serves the table from its own mmap'd chunk, or the size field of the next heap chunk taken as a
    double (a denormal around 1e-318) when an earlier free has raised the mmap threshold: nothing in either
    case for any sum the matchers form."""
    from oracle import pyoracle
    if not (pyoracle.available("reference") or os.path.isdir("/root/reference")):
        pytest.skip("needs the compiled reference")
    ref = pyoracle.load("reference")
    assert ref.value_probability(0) == 0.0
    assert ref.value_probability(1) == 1e-3
    assert abs(ref.value_probability(65534) - (1e-3 + 0.998 * 65533 / 65534)) < 1e-15
    assert 0.0 <= ref.value_probability(65535) < 1e-300
