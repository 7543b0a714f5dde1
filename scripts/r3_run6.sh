#!/bin/bash
out=gpurun_out/r3_run6.log; : > $out
( timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "graphed_hot_path" 2>&1 | tail -6 ) >> $out
timeout 300 python scripts/time_small_batch.py > gpurun_out/r3_small_batch.json 2>> $out
cat gpurun_out/r3_small_batch.json >> $out
cat $out
