"""In-tree build of libcsm_b200.so (nvcc, sm_100a only)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libcsm_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "csm_b200.cu")]
HEADERS = [os.path.join(HERE, "csrc", "csm_kernels.cuh"),
           os.path.join(HERE, "csrc", "csm_device.cuh"),
           os.path.join(HERE, "csrc", "csm_window_tma.cuh"),
           os.path.join(HERE, "csrc", "csm_refine.cuh"),
           os.path.join(HERE, "csrc", "csm_bounds.cuh"),
           os.path.join(HERE, "csrc", "csm_mapbuild.cuh"),
           os.path.join(ROOT, "include", "csm_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--cudart", "shared", "-ldl",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA extension for sm_100a. Returns the path of the .so."""
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + [
        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"),
        "-o", LIB] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
