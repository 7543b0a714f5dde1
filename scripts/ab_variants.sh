#!/bin/bash
# A/B timing of library variants built by scripts/build_variants.py (GPU box).
#   scripts/ab_variants.sh name1 name2 ...   -> gpurun_out/ab_variants.log
out=gpurun_out/ab_variants.log; : > $out
V=cadence_gemma_b200/csrc/variants
for v in "$@"; do
  echo "== $v cfg2" >> $out
  CG_B200_LIB=$PWD/$V/lib_$v.so timeout 120 python scripts/fused_check.py --case cfg2 >> $out 2>&1
done
for rep in 1 2 3; do
  for v in "$@"; do
    echo "== $v time" >> $out
    CG_B200_LIB=$PWD/$V/lib_$v.so timeout 120 python scripts/fused_check.py --case time 2>&1 | tail -1 >> $out
  done
done
python - <<'P' >> $out
import json,re,collections
d=collections.defaultdict(list); cur=None
for l in open("gpurun_out/ab_variants.log"):
    m=re.match(r"== (\S+) time",l)
    if m: cur=m.group(1); continue
    if l.startswith("{") and '"time"' in l and cur: d[cur].append(json.loads(l)["us_median"]); cur=None
for k,v in d.items(): print("SUMMARY",k,[round(x,1) for x in v])
P
grep "SUMMARY\|identical\|watchdog\": [1-9]" $out | cut -c1-300
