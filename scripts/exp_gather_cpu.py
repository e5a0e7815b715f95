"""Builds and runs scripts/exp_gather_cpu.cpp with the BlockGatherer class cut out of the product source."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "my_lidar_graph_slam_v2_b200", "host", "src", "loop_detector.cpp")).read()
a = src.index("class BlockGatherer")
b = src.index("namespace {", a)
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "gatherer_class.inc"), "w").write(src[a:b])
    exe = os.path.join(d, "gather")
    subprocess.run(["g++", "-O3", "-std=c++17", "-msse4.1", "-pthread", "-I", d,
                    os.path.join(ROOT, "scripts", "exp_gather_cpu.cpp"), "-o", exe], check=True)
    subprocess.run([exe] + sys.argv[1:], check=True)
