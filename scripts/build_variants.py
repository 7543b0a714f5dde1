"""Builds tuning variants of libcadence_b200.so with extra -D switches.

    python scripts/build_variants.py name:-DCGF_LOOK=1 name2:-DCGF_SLEEP_NS=0,-DCGF_PRELOAD=0

Outputs cadence_gemma_b200/csrc/variants/lib_<name>.so; select one at run time
with CG_B200_LIB=<path>.
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadence_gemma_b200 import build as B  # noqa: E402

out_dir = os.path.join(B.CSRC, "variants")
os.makedirs(out_dir, exist_ok=True)
procs = []
for spec in sys.argv[1:]:
  name, _, defs = spec.partition(":")
  out = os.path.join(out_dir, f"lib_{name}.so")
  cmd = [B.find_nvcc(), "-ccbin", "/usr/bin/g++", *B.NVCC_FLAGS, *[d for d in defs.split(",") if d],
         "-o", out, os.path.join(B.CSRC, "cadence_b200.cu")]
  procs.append((out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for out, pr in procs:
  log, _ = pr.communicate()
  if pr.returncode != 0:
    print(log)
    raise SystemExit(f"build of {out} failed")
  print(out)
