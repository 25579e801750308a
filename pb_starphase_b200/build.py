"""Builds libstarphase_gpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libstarphase_gpu.so"
SOURCES = ["starphase_gpu.cu", "sp_kernels.cuh", "sp_align.cuh", "sp_addchain.inc", "../../include/starphase_gpu.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; cannot build libstarphase_gpu.so")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / s).resolve().stat().st_mtime > t for s in SOURCES)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    subprocess.check_call([sys.executable, str(CSRC / "gen_addchain.py")], cwd=CSRC)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "starphase_gpu.cu")]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


# ---------------------------------------------------------------------------------------------
# C++ host above the C ABI (pb_starphase_b200/host): libstarphase_host.so + the pybind11 test module
# ---------------------------------------------------------------------------------------------
HOST = HERE / "host"
HOST_LIB = HERE / "libstarphase_host.so"
HOST_SOURCES = ["sp_host_core.cpp", "sp_host_gpu.cpp", "sp_host_hla.cpp", "sp_host_cyp2d6.cpp", "sp_host_debug.cpp"]


def host_module_path() -> Path:
    import sysconfig

    return HERE / ("_starphase_host" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_host(force: bool = False) -> Path:
    """g++ build of the C++ host library and its pybind11 module, both in-tree, linked against libstarphase_gpu.so
    through $ORIGIN so the three files travel together."""
    import pybind11
    import sysconfig

    mod = host_module_path()
    deps = [HOST / s for s in HOST_SOURCES + ["sp_host_py.cpp", "starphase_host.hpp"]] + [HERE.parent / "include" / "starphase_gpu.h"]
    newest = max(d.stat().st_mtime for d in deps)
    if not force and HOST_LIB.exists() and mod.exists() and min(HOST_LIB.stat().st_mtime, mod.stat().st_mtime) > newest:
        return mod
    gxx = shutil.which("g++") or "/usr/bin/g++"
    common = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", "-Wl,-rpath,$ORIGIN", "-L", str(HERE)]
    subprocess.check_call(common + ["-o", str(HOST_LIB)] + [str(HOST / s) for s in HOST_SOURCES] + ["-lstarphase_gpu"])
    subprocess.check_call(common + ["-fvisibility=hidden", "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
                                    "-o", str(mod), str(HOST / "sp_host_py.cpp"), "-lstarphase_host", "-lstarphase_gpu"])
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
