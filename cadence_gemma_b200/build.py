"""Builds the sm_100a shared library in-tree with nvcc (cross-compiles on CPU).

    python -m cadence_gemma_b200.build [--force] [--verbose]

Output: cadence_gemma_b200/csrc/libcadence_b200.so (git-ignored, travels to the
GPU box with the repo snapshot).  Links the CUDA runtime statically, so the
library has no load-time dependency on libcuda / libcudart.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcadence_b200.so")
SOURCES = ["cadence_b200.cu"]
HEADERS = ["cg_common.cuh", "cg_scan.cuh", "cg_conv1d.cuh", "cg_fused.cuh", "cg_decode.cuh", "cg_train.cuh",
           os.path.join("..", "..", "include", "cadence_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def find_nvcc() -> str:
  for cand in (os.environ.get("NVCC"), shutil.which("nvcc"),
               "/usr/local/cuda/bin/nvcc"):
    if cand and os.path.exists(cand):
      return cand
  raise RuntimeError("nvcc not found; cannot build libcadence_b200.so")


def is_stale() -> bool:
  if not os.path.exists(LIB):
    return True
  built = os.path.getmtime(LIB)
  deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
  return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
  if not force and not is_stale():
    return LIB
  cmd = [find_nvcc(), *NVCC_FLAGS]
  if verbose:
    cmd += ["-Xptxas", "-v"]
  cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
  env = dict(os.environ)
  # the image exports CC/CXX=/opt/gcc wrappers; nvcc wants the system g++
  if os.path.exists("/usr/bin/g++"):
    cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
  proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
  if verbose or proc.returncode != 0:
    sys.stderr.write(proc.stdout + proc.stderr)
  if proc.returncode != 0:
    raise RuntimeError("nvcc failed building libcadence_b200.so")
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
