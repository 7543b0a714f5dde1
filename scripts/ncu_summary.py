"""Summarises an .ncu-rep: key metrics, opcode mix per element, stall reasons.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--elems 41943040]
"""
import csv
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
elems = 41943040
if "--elems" in sys.argv:
  elems = int(sys.argv[sys.argv.index("--elems") + 1])

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for h, u, v in zip(hdr, units, vals):
  if h in want:
    print(f"  {h:72s} {v:>16s} {u}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
byop, stalls = Counter(), Counter()
total = 0
for r in rows[2:]:
  if len(r) < len(hdr):
    continue
  m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[idx["Source"]])
  op = m.group(2) if m else r[idx["Source"]][:12]
  n = int(r[idx["Instructions Executed"]] or 0)
  total += n
  byop[op] += n
  for c in hdr:
    if c.startswith("stall_") and "Not Issued" not in c and r[idx[c]]:
      stalls[c] += int(r[idx[c]])
print(f"warp instructions {total}  = {total * 32 / elems:.2f} thread-instr / element")
print("opcode mix (thread-instr / element):")
line = []
for op, n in byop.most_common(36):
  line.append(f"{op} {n * 32 / elems:.2f}")
for i in range(0, len(line), 6):
  print("   " + " | ".join(line[i:i + 6]))
tot = sum(stalls.values())
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in stalls.most_common(10)))
