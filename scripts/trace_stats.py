"""Per-phase mean cycles from gpurun_out/fused_trace.json (scripts/fused_trace.py)."""
import json
import statistics
import sys

r = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fused_trace.json"))
for wg in ["wg0", "wg1", "wg2", "wg3"]:
  tiles, cur = [], {}
  for t, c in r[wg]:
    cur[c] = t
    if c == 6:
      tiles.append(cur)
      cur = {}
  d = lambda a, b: statistics.mean(x[b] - x[a] for x in tiles[2:] if a in x and b in x)
  per = statistics.mean(tiles[i + 1][1] - tiles[i][1] for i in range(2, len(tiles) - 1))
  print(f"{wg}: wait t_full {d(1,2):6.0f} accload {d(2,3):6.0f} gate {d(3,4):6.0f} lookback {d(4,5):6.0f} "
        f"pass2 {d(5,6):6.0f} period {per:6.0f} tiles {len(tiles)}")
tiles, cur = [], {}
for t, c in r["mma"]:
  cur[c] = t
  if c == 4:
    tiles.append(cur)
    cur = {}
d = lambda a, b: statistics.mean(x[b] - x[a] for x in tiles[2:])
per = statistics.mean(tiles[i + 1][1] - tiles[i][1] for i in range(2, len(tiles) - 1))
print(f"mma: wait x_full {d(1,2):6.0f} wait t_empty {d(2,3):6.0f} issue {d(3,4):6.0f} period {per:6.0f}")
tiles, cur = [], {}
for t, c in r["producer"]:
  cur[c] = t
  if c == 2:
    tiles.append(cur)
    cur = {}
print("producer wait x_empty", statistics.mean(x[2] - x[1] for x in tiles[2:]), "end", r["wg0"][-1][0])
