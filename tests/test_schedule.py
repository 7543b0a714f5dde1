"""Work schedule of the fused tcgen05 kernels (cg::fused::Schedule, cadence_gemma_b200/csrc/cg_fused.cuh):
every MMA tile of every family is assigned to exactly one CTA, with and without tail stealing, for every
grid size from 1 to 160 CTAs -- a host-side sweep compiled from tests/native/schedule_coverage.cu (no GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not found")
def test_schedule_covers_every_tile_exactly_once(tmp_path):
  nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
  exe = str(tmp_path / "schedule_coverage")
  cmd = [nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w",
         "-I", os.path.join(ROOT, "cadence_gemma_b200", "csrc"), "-o", exe,
         os.path.join(ROOT, "tests", "native", "schedule_coverage.cu")]
  if os.path.exists("/usr/bin/g++"):
    cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
  subprocess.run(cmd, check=True, capture_output=True, text=True)
  out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
  assert "sweep ok rc=0" in out, out
  # the balance the stealing is for: config 2 (148 CTAs, 20 families, 256 MMA tiles per family)
  assert "ok G=148 nfam=20 npairs=256 steal=0 load 32..37" in out
  assert "ok G=148 nfam=20 npairs=256 steal=11 load 33..35" in out
