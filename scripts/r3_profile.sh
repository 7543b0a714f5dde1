#!/bin/bash
# round 2, second session: profiling run (B200_PROFILING.md recipe): plain runs first, then ncu on the SAME commands
out=gpurun_out/r3_profile.log; : > $out
export CG_BENCH_RAMP_MS=0
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
P1="python scripts/profile_fused.py --B 8 --T 2048 --conv 1"
$B > gpurun_out/r3_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_bench_launches_ncu.csv $B > gpurun_out/r3_ncu_bench.log 2>&1
echo "launch list rc=$?" >> $out
$P1 > gpurun_out/r3_plain_p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rglru_fused -s 2 -c 1 -f -o gpurun_out/r3_prof_fused_conv $P1 > gpurun_out/r3_ncu_p1.log 2>&1
echo "one-launch capture rc=$?" >> $out
ls -la gpurun_out/r3_prof_* >> $out 2>&1
cat $out; tail -3 gpurun_out/r3_ncu_p1.log
