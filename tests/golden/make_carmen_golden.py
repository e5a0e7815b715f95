"""Golden vectors for the Carmen log reader: a log that holds every record type the reference reads
(io/carmen/carmen_reader.cpp) and what the COMPILED REFERENCE's CarmenLogReader::Load makes of it
(oracle/_ref/libcsm_ref.so). Run in the build container (needs /root/reference for the checker build):

    python tests/golden/make_carmen_golden.py

Doubles are stored as hex strings (float.hex), so the comparison in tests/test_carmen_io.py is exact."""
import hashlib
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "carmen_vectors.json")


def make_log(seed=7, with_params=True):
    rng = np.random.default_rng(seed)
    g = lambda v: "%.17g" % v
    rs = lambda n: " ".join(g(v) for v in rng.uniform(0.2, 30.0, n))
    lines = ["# message_name [message contents] ipc_timestamp ipc_hostname logger_timestamp"]
    # old-format scans before any PARAM: the reader guesses the beam geometry from the beam count
    for n in (180, 181, 360, 361, 400, 401, 77):
        lines.append("FLASER %d %s 1.25 -0.5 0.3 1.0 -0.25 0.25 %s host 0.5" % (n, rs(n), g(100.0 + n)))
    lines.append("LASER3 181 " + rs(181))
    if with_params:
        lines += ["PARAM Laser.MinRange 0.1", "PARAM Laser.MaxRange 25.5", "PARAM Laser.AngleIncrement 0.0174",
                  "PARAM Laser.MinAngle -1.5", "PARAM Laser.MinRange 9.9",           # the first value stays
                  "PARAM robot_length 0.6", "PARAM onlyname"]
    lines.append("ODOM 1.5 -2.25 0.785398 0.4 0.05 0.01 12.5 robothost 12.625")
    lines.append("TRUEPOS 1 2 3 4 5 6 7 host 8")
    lines.append("RAWLASER1 0 -1.5707963267948966 3.1415926535897931 0.017453292519943295 81.9 0.01 0 90 %s 3 1 2 3 "
                 "20.25 host 20.5" % rs(90))
    lines.append("RAWLASER3 0 -0.5 1.0 0.05 30 0.01 1 20 %s 0 21.0 host 21.5" % rs(20))
    lines.append("ROBOTLASER1 0 -3.1415926535897931 6.2831853071795862 0.017453292519943295 40.0 0.01 0 360 %s "
                 "10.35 4.1 0.52 10.0 4.0 0.5 0.3 0.02 0.5 0.4 1e30 31.0 host 31.25" % rs(360))
    lines.append("ROBOTLASER2 0 -1.0 2.0 0.1 12.5 0.01 0 21 %s -3.0 2.0 -2.5 -3.2 2.1 -2.4 0.1 0.0 0.5 0.4 0 "
                 "32.0 host 32.25" % rs(21))
    lines.append("FLASER 181 %s 5.5 6.5 0.75 5.25 6.25 0.7 41.0 host 41.5" % rs(181))
    lines.append("RLASER 90 %s 5.0 6.0 3.9 5.25 6.25 0.7 42.0 host 42.5" % rs(90))
    lines.append("LASER4 33 " + rs(33))
    lines.append("NMEAGGA 1 2 3 4")
    lines.append("ODOM -7 8 -3.0 0 0 0 50.0 host 50.0")
    return "\n".join(lines) + "\n"


def encode(records):
    out = []
    for r in records:
        e = {}
        for k, v in r.items():
            if isinstance(v, str):
                e[k] = v
            elif np.ndim(v) == 0:
                e[k] = float(v).hex()
            elif len(v) <= 3:
                e[k] = [float(x).hex() for x in v]
            else:       # beam arrays: length, first, last and the SHA-256 of the packed doubles
                a = np.ascontiguousarray(v, dtype="<f8")
                e[k] = {"n": len(a), "first": float(a[0]).hex(), "last": float(a[-1]).hex(),
                        "sha256": hashlib.sha256(a.tobytes()).hexdigest()}
        out.append(e)
    return out


def main():
    from oracle import pyoracle
    from my_lidar_graph_slam_v2_b200 import hostapi
    ref = pyoracle.load("reference")
    cases = []
    for seed, with_params in ((7, True), (8, False)):
        text = make_log(seed, with_params)
        h = ref.carmen_load(text)
        recs = hostapi.carmen_records(ref.lib, "orc_carmen_", h)
        ref.lib.orc_carmen_destroy(h)
        cases.append({"seed": seed, "with_params": with_params, "log": text, "records": encode(recs)})
    metrics = {"int": [0.0, 1.9, 1234567.2, 86.5], "float": [0.1, 2.0 / 3.0, 1234.56789, 1e-7, 3.0e6]}
    strings = {"int": ref.metric_values_string(0, metrics["int"]), "float": ref.metric_values_string(1, metrics["float"]),
               "uint64": ref.metric_values_string(2, [0.0, 4096.0, 3.0e9])}
    with open(OUT, "w") as f:
        json.dump({"generator": "tests/golden/make_carmen_golden.py", "cases": cases,
                   "metric_values": metrics, "metric_strings": strings}, f, indent=0)
    print("wrote", OUT, [len(c["records"]) for c in cases], strings)


if __name__ == "__main__":
    main()
