#!/bin/bash
# (gpurun brings back at most 64 MiB: two full captures of ~23 MB each per call)
# round 2 profiling run (B200_PROFILING.md recipe): plain runs first, then ncu on the SAME commands
out=gpurun_out/r2_profile.log; : > $out
export CG_BENCH_RAMP_MS=0
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
P0="python scripts/profile_fused.py --B 8 --T 2048 --conv 0"
P1="python scripts/profile_fused.py --B 2 --T 2048 --conv 1"
$B > gpurun_out/r2_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?" >> $out
$P0 > gpurun_out/r2_plain_p0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rglru_fused -s 2 -c 1 -f -o gpurun_out/r2_prof_fused $P0 > gpurun_out/r2_ncu_p0.log 2>&1
echo "fused capture rc=$?" >> $out
$P1 > gpurun_out/r2_plain_p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rglru_fused -s 2 -c 1 -f -o gpurun_out/r2_prof_fused_conv $P1 > gpurun_out/r2_ncu_p1.log 2>&1
echo "fused+conv capture rc=$?" >> $out
ls -la gpurun_out/r2_prof_* >> $out 2>&1
cat $out; tail -3 gpurun_out/r2_ncu_p0.log
