"""The on-disk formats either side of the full loop: the C++ CarmenLogReader /
CarmenLogWriter / WriteMetricsJson of host/ against the compiled reference's CarmenLogReader
(io/carmen/carmen_reader.cpp) and its value-sequence strings (metric/metric.hpp), live and through the
committed golden vectors (tests/golden/carmen_vectors.json, made by make_carmen_golden.py)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from my_lidar_graph_slam_v2_b200 import hostapi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_carmen_golden import encode, make_log  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "carmen_vectors.json")))


def host_records(text):
    log = hostapi.CarmenLog(text=text)
    try:
        return log.records()
    finally:
        log.close()


def reference():
    from oracle import pyoracle
    if not pyoracle.available("reference"):
        pytest.skip("compiled reference not built")
    ref = pyoracle.load("reference")
    if ref.carmen_load("") is None:
        pytest.skip("checker predates the Carmen reader")
    return ref


@pytest.mark.parametrize("case", range(len(GOLDEN["cases"])))
def test_reader_matches_reference_golden(case):
    g = GOLDEN["cases"][case]
    assert make_log(g["seed"], g["with_params"]) == g["log"]          # the generator is deterministic
    got = encode(host_records(g["log"]))
    assert len(got) == len(g["records"]) == 17
    for a, b in zip(got, g["records"]):
        assert a == b                                                  # every double bit for bit (hex / SHA-256)


@pytest.mark.parametrize("seed,with_params", [(7, True), (8, False), (21, True), (22, False)])
def test_reader_matches_compiled_reference(seed, with_params):
    ref = reference()
    text = make_log(seed, with_params)
    h = ref.carmen_load(text)
    want = hostapi.carmen_records(ref.lib, "orc_carmen_", h)
    ref.lib.orc_carmen_destroy(h)
    got = host_records(text)
    assert [r["sensor_id"] for r in got] == [r["sensor_id"] for r in want]
    for a, b in zip(got, want):
        assert a.keys() == b.keys()
        for k in a:
            if isinstance(a[k], str):
                assert a[k] == b[k]
            else:
                assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (a["sensor_id"], k)


def test_reader_number_formats_match_compiled_reference():
    """tabs, runs of blanks, exponents, signs, trailing blanks, CR line ends, integers written as reals' prefixes"""
    ref = reference()
    rng = np.random.default_rng(11)
    def num(v):
        return rng.choice(["%.17g" % v, "%.6e" % v, "%+.9f" % v, "%.3f" % v])
    lines = []
    for k in range(40):
        n = int(rng.integers(3, 12))
        sep = rng.choice([" ", "  ", "\t", " \t "])
        r = sep.join(num(v) for v in rng.uniform(0.1, 40.0, n))
        kind = k % 4
        if kind == 0:
            body = ["ROBOTLASER1", "0", num(-1.5), num(3.0), num(3.0 / n), num(30.0), "0.01", "0", str(n), r,
                    num(1.0), num(2.0), num(0.5), num(0.9), num(1.9), num(0.45), "0", "0", "0", "0", "0",
                    num(100.0 + k), "host", num(100.5 + k)]
        elif kind == 1:
            body = ["FLASER", str(n), r, num(1.0), num(-2.0), num(-0.5), num(1.1), num(-2.1), num(-0.6),
                    num(200.0 + k), "h", num(200.0 + k)]
        elif kind == 2:
            body = ["ODOM", num(rng.normal()), num(rng.normal()), num(rng.normal()), num(0.3), num(-0.1), "0",
                    num(300.0 + k), "h", num(300.0 + k)]
        else:
            body = ["RAWLASER2", "0", num(-0.7), num(1.4), num(1.4 / n), num(20.0), "0.01", "1", str(n), r, "2", "7", "9",
                    num(400.0 + k), "h", num(400.0 + k)]
        lines.append(sep.join(body) + rng.choice(["", " ", "\r", "  \t"]))
    text = "\n".join(lines) + "\n"
    h = ref.carmen_load(text)
    want = hostapi.carmen_records(ref.lib, "orc_carmen_", h)
    ref.lib.orc_carmen_destroy(h)
    got = host_records(text)
    assert len(got) == len(want) == 40
    assert encode(got) == encode(want)


def test_reader_edge_cases():
    assert host_records("") == []
    assert host_records("\n\n# only comments\nUNKNOWN 1 2 3\n") == []
    # a record cut short keeps zeros, beams that are missing read as 0
    (r,) = host_records("FLASER 4 1.0 2.0\n")
    assert r["kind"] == "scan" and list(r["ranges"]) == [1.0, 2.0, 0.0, 0.0] and r["time_stamp"] == 0.0
    (r,) = host_records("ODOM 1 2\n")
    assert list(r["odom_pose"]) == [1.0, 2.0, 0.0]
    with pytest.raises(FileNotFoundError):
        hostapi.CarmenLog(path="/nonexistent/log.carmen")


@pytest.mark.parametrize("old_format", [False, True])
def test_writer_round_trip_is_exact(tmp_path, old_format):
    rng = np.random.default_rng(5)
    n_scans, n_beams = 9, 360
    ranges = rng.uniform(0.05, 29.0, (n_scans, n_beams))
    poses = np.cumsum(rng.normal(0, 0.3, (n_scans, 3)), axis=0)
    stamps = 1e9 + np.arange(n_scans) * 0.1 + rng.uniform(0, 1e-3, n_scans)
    start, inc, rmax = -np.pi, 2 * np.pi / n_beams, 30.0
    rel = (0.12, -0.03, 0.25)
    path = str(tmp_path / "run.log")
    hostapi.write_carmen_log(path, ranges, poses, stamps, start, inc, rmax, laser_on_robot=rel, old_format=old_format)
    log = hostapi.CarmenLog(path=path)
    recs = log.records()
    log.close()
    assert [r["kind"] for r in recs] == ["odom", "scan"] * n_scans
    scans = [r for r in recs if r["kind"] == "scan"]
    for k, r in enumerate(scans):
        assert np.array_equal(r["ranges"], ranges[k])
        assert np.array_equal(r["odom_pose"], poses[k]) and r["time_stamp"] == stamps[k]
        assert np.array_equal(r["angles"], start + inc * np.arange(n_beams))
        assert r["max_range"] == rmax and r["min_range"] == 0.0
        # the sensor pose on the robot comes back through Compound / InverseCompound: a few ulp
        assert np.allclose(r["relative_sensor_pose"], rel, rtol=0, atol=1e-12)
    # the reference reads the written file the same way
    from oracle import pyoracle
    if pyoracle.available("reference"):
        ref = pyoracle.load("reference")
        h = ref.carmen_load(open(path).read())
        if h is not None:
            want = hostapi.carmen_records(ref.lib, "orc_carmen_", h)
            ref.lib.orc_carmen_destroy(h)
            assert encode(want) == encode(recs)


def test_metric_value_strings_match_reference():
    m, s = GOLDEN["metric_values"], GOLDEN["metric_strings"]
    assert hostapi.metric_values_string("Frontend.ProcessTime", m["int"]) == s["int"]
    assert hostapi.metric_values_string("LocalSlam.ScanMatcherCorrelative.ScoreValue", m["float"]) == s["float"]
    assert hostapi.metric_values_string("X.PrecompMapMemoryUsage", [0.0, 4096.0, 3.0e9]) == s["uint64"]
    ref = reference()
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.uniform(0, 1e6, 50), rng.uniform(0, 1, 50), [0.0, 0.5, 1.5, 2.5, 1e-7]])
    assert hostapi.metric_values_string("Backend.LoopDetectionTime", v) == ref.metric_values_string(0, v)
    assert hostapi.metric_values_string("Frontend.IntervalTravelDist", v) == ref.metric_values_string(1, v)


def test_metrics_json_layout():
    # slam_launcher.cpp:171-181 + metric.cpp:460-496: five families, ids as keys (dots are not paths), every
    # leaf a string
    seqs = {"Frontend.ProcessTime": [120.7, 98.2], "LoopDetector.BranchBound.ScoreValue": [0.625],
            'odd "id"/x': [1.0]}
    text = hostapi.metrics_json(seqs)
    doc = json.loads(text)
    assert list(doc) == ["Counters", "Gauges", "Distributions", "Histograms", "ValueSequences"]
    assert all(doc[k] == "" for k in list(doc)[:4])
    vs = doc["ValueSequences"]
    assert vs["Frontend.ProcessTime"] == {"NumOfSamples": "2", "Values": "120 98"}
    assert vs["LoopDetector.BranchBound.ScoreValue"] == {"NumOfSamples": "1", "Values": "0.625000"}
    assert vs['odd "id"/x']["Values"] == "1"
    assert json.loads(hostapi.metrics_json({}))["ValueSequences"] == ""
