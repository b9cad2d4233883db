"""Builds the in-tree native libraries.

  navigation2_b200/libmppi_b200.so   the product: CUDA kernels + C ABI (include/mppi_b200.h), sm_100a only
  navigation2_b200/libmppi_config.so the ABI's host-only configuration helpers on their own (no CUDA)
  oracle/_build/*.so, oracle/_ref/   the CPU oracle (test infrastructure) via oracle/Makefile

`python -m navigation2_b200.build` builds both; `__graft_entry__.build()` calls build_all().
nvcc cross-compiles sm_100a without a GPU; the resulting .so files travel to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "navigation2_b200", "csrc")
LIB = os.path.join(ROOT, "navigation2_b200", "libmppi_b200.so")

SOURCES = ["mppi_kernels.cu", "mppi_capi.cu", "mppi_config.cpp", "costmap_inflation.cu"]
HEADERS = [
    os.path.join(ROOT, "include", "mppi_b200.h"),
    os.path.join(ROOT, "include", "mppi_math.h"),
    os.path.join(ROOT, "include", "costmap_inflation_b200.h"),
    os.path.join(CSRC, "mppi_device.h"),
    os.path.join(CSRC, "mppi_kernels.h"),
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # no FMA contraction on the device or the host side: bit-identical floats with the CPU oracle
    # on everything that feeds a costmap index
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-fopenmp",
    "-shared", "-lgomp",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the product is CUDA-only, there is no CPU build")
    return nvcc


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_product(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if not force and not _stale(LIB, srcs + HEADERS + [os.path.abspath(__file__)]):
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, *srcs, "-o", LIB]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=ROOT)
    return LIB


CONFIG_LIB = os.path.join(ROOT, "navigation2_b200", "libmppi_config.so")


def build_config(force: bool = False, verbose: bool = False) -> str:
    """csrc/mppi_config.cpp on its own (plain C++, no CUDA): what `params` needs to describe a workload, so the
    reference arm of bench.py and the oracle-only tests do not map the product library."""
    src = os.path.join(CSRC, "mppi_config.cpp")
    if not force and not _stale(CONFIG_LIB, [src, os.path.join(ROOT, "include", "mppi_b200.h")]):
        return CONFIG_LIB
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), src, "-o", CONFIG_LIB]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=ROOT)
    return CONFIG_LIB


HOST = os.path.join(ROOT, "navigation2_b200", "host")
HOST_LIB = os.path.join(ROOT, "navigation2_b200", "libmppi_controller_b200.so")
HOST_TEST = os.path.join(HOST, "test_host")


def build_host(force: bool = False, verbose: bool = False) -> str:
    """The C++ host mirror of the reference's plugin classes (navigation2_b200/host) on top of the C ABI,
    and its test driver.  Built against the in-tree ROS type shim; with a ROS 2 underlay it is built by colcon."""
    src = os.path.join(HOST, "src", "mppi.cpp")
    hdrs = [os.path.join(HOST, "include", "nav2_mppi_controller_b200", "mppi.hpp"),
            os.path.join(HOST, "include", "ros_shim", "ros_shim.hpp"), os.path.join(ROOT, "include", "mppi_b200.h"),
            os.path.join(HOST, "include", "nav2_costmap_2d_b200", "inflation_layer.hpp"),
            os.path.join(HOST, "include", "nav2_controller_b200", "feasible_path_handler.hpp"),
            os.path.join(HOST, "include", "nav2_controller_b200", "fleet_controller_server.hpp"),
            os.path.join(ROOT, "include", "costmap_inflation_b200.h")]
    test_src = os.path.join(HOST, "test", "test_host.cpp")
    inc = ["-I", os.path.join(HOST, "include"), "-I", os.path.join(ROOT, "include")]
    common = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-fPIC", *inc]
    link = ["-L", os.path.dirname(LIB), "-lmppi_b200", "-Wl,-rpath,$ORIGIN"]
    if force or _stale(HOST_LIB, [src, LIB, *hdrs]):
        cmd = [*common, "-shared", src, *link, "-o", HOST_LIB]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=ROOT)
    if force or _stale(HOST_TEST, [test_src, HOST_LIB, *hdrs]):
        cmd = [*common, test_src, "-L", os.path.dirname(LIB), "-lmppi_controller_b200", "-lmppi_b200",
               "-Wl,-rpath,$ORIGIN/..", "-o", HOST_TEST]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=ROOT)
    return HOST_LIB


def build_oracle(verbose: bool = False) -> None:
    """Builds the checker (oracle/) and, when /root/reference exists, oracle/_ref from the reference's own header."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


def build_all(force: bool = False, verbose: bool = False) -> None:
    # one builder at a time (torchrun starts several ranks in the same tree)
    import fcntl
    with open(os.path.join(ROOT, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            build_product(force=force, verbose=verbose)
            build_config(force=force, verbose=verbose)
            build_host(force=force, verbose=verbose)
            build_oracle(verbose=verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
