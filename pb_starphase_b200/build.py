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
# translation units of the library and the headers each one depends on
HEADERS = ["sp_internal.cuh", "sp_kernels.cuh", "sp_addchain.inc", "../../include/starphase_gpu.h"]
UNITS = {
    "starphase_gpu.cu": HEADERS + ["sp_misc.cuh"],   # context, K1 / K2 / K3 / K5 / K6
    "sp_align.cu": HEADERS + ["sp_align.cuh"],       # K4
    "sp_comm.cu": HEADERS + ["sp_comm_kernels.cuh"],  # multi-GPU (NCCL, loaded with dlopen at run time)
    "sp_consensus.cu": HEADERS + ["sp_consensus.cuh"],  # K7
    "sp_graph.cu": HEADERS + ["sp_graph.cuh"],  # K8
    "sp_affine.cu": HEADERS + ["sp_affine.cuh", "sp_align.cuh"],  # K9
}
OBJDIR = CSRC / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; cannot build libstarphase_gpu.so")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any((CSRC / d).resolve().stat().st_mtime > t for d in deps)


def needs_build() -> bool:
    return any(_stale(LIB, [u] + deps) for u, deps in UNITS.items())


def build(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -c per translation unit (objects cached under csrc/build/, rebuilt when the unit or one of its headers changed),
    then one shared-library link with the static CUDA runtime."""
    if not force and not needs_build():
        return LIB
    subprocess.check_call([sys.executable, str(CSRC / "gen_addchain.py")], cwd=CSRC)
    OBJDIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for unit, deps in UNITS.items():
        obj = OBJDIR / (Path(unit).stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [unit] + deps):
            cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(CSRC / unit)]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
                print(" ".join(cmd))
            procs.append((unit, subprocess.Popen(cmd, cwd=CSRC)))
    for unit, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, f"nvcc {unit}")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", str(LIB),
                           *map(str, objs), "-ldl"], cwd=CSRC)
    return LIB


# ---------------------------------------------------------------------------------------------
# C++ host above the C ABI (pb_starphase_b200/host): libstarphase_host.so + the pybind11 test module
# ---------------------------------------------------------------------------------------------
HOST = HERE / "host"
HOST_LIB = HERE / "libstarphase_host.so"
HOST_SOURCES = ["sp_host_core.cpp", "sp_host_gpu.cpp", "sp_host_hla.cpp", "sp_host_cyp2d6.cpp", "sp_host_debug.cpp", "sp_host_consensus.cpp", "sp_host_graph.cpp"]


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
