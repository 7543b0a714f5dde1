#!/bin/bash
# round 2, GPU run 2: the whole GPU suite + smoke + bench (own and reference arm)
out=gpurun_out/r2_run2.log; : > $out
( timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 ) >> $out
echo "== smoke" >> $out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 ) >> $out
echo "== bench" >> $out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?" >> $out
echo "== bench reference" >> $out
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "rc=$?" >> $out
tail -c 6000 $out
cut -c1-3000 gpurun_out/r2_bench_n1.json
tail -5 gpurun_out/r2_bench_n1.err
