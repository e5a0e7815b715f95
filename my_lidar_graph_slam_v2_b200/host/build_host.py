"""In-tree build of libcsm_host.so: the C++ plugin mirror on top of libcsm_b200.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libcsm_host.so")
SOURCES = [os.path.join(HERE, "src", f) for f in
           ("cost_square_error.cpp", "scan_matchers.cpp", "loop_detector.cpp", "loop_searcher.cpp", "map_builder.cpp", "slam_pipeline.cpp", "carmen_log.cpp", "c_shim.cpp")]


def build():
    cmd = ["g++", "-std=c++17", "-O3", "-g", "-rdynamic", "-ffp-contract=off", "-msse4.1", "-fPIC", "-shared", "-Wall",
           "-I", os.path.join(HERE, "include"), "-I", os.path.join(ROOT, "include"),
           "-o", LIB] + SOURCES + ["-L", PKG, "-l:libcsm_b200.so", "-Wl,-rpath,$ORIGIN", "-pthread"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build())
