"""Build libroadsurf_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch arch list)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.environ.get("ROADSURF_B200_LIB") or os.path.join(LIB_DIR, "libroadsurf_b200.so")   # env: an experiment build
SOURCES = ["rs_zonal.cu", "rs_tables.cu", "rs_extract.cu", "rs_pairs.cu", "rs_api.cu", "rs_comm.cu", "rs_wide.cu", "rs_fstats.cu", "rs_codec.cu",
           "rs_hostcopy.cu"]
HEADERS = [os.path.join(CSRC, "rs_internal.h"), os.path.join(CSRC, "rs_raster.cuh"), os.path.join(CSRC, "rs_codec_core.h"),
           os.path.join(ROOT, "include", "roadsurf_b200.h")]

# -fmad=false: the rounding of every float64 operation of the rasterizer is part of the
# specification (GDAL evaluates a*b+c with separate multiply and add on x86-64).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-fmad=false", "-t", "0",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libroadsurf_b200.so cannot be built")
    return p


def is_stale() -> bool:
    if os.environ.get("ROADSURF_B200_LIB"):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB_PATH]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("RS_NVCC_EXTRA", "").split()
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl", "-lpthread"]
    subprocess.check_call(cmd)
    return LIB_PATH
