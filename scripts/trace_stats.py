"""Mean cycles between consecutive pipeline events of gpurun_out/fused_trace.json
(scripts/fused_trace.py), per role, keyed by (previous event -> next event)."""
import collections
import json
import sys

r = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/fused_trace.json"))
for role, ev in r.items():
  ev = ev[len(ev) // 6:len(ev) - len(ev) // 6]          # steady state only
  if len(ev) < 4:
    continue
  acc = collections.defaultdict(list)
  for (t0, c0), (t1, c1) in zip(ev, ev[1:]):
    acc[(c0, c1)].append(t1 - t0)
  first = [t for t, c in ev if c == ev[0][1]]
  period = (first[-1] - first[0]) / max(1, len(first) - 1)
  parts = "  ".join(f"{a}->{b}: {sum(v) / len(v):6.0f} (n={len(v)})" for (a, b), v in sorted(acc.items()))
  print(f"{role:9s} period {period:7.0f}  {parts}")
print("end", max(ev[-1][0] for ev in r.values() if ev), "cycles; events per role", {k: len(v) for k, v in r.items()})
