#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_host_shim.py -x -q -m gpu -k "sparse or overflow or pairlist or c3 or C3 or rows or cell or droplet or smp" > $out/tests.log 2>&1; tail -3 $out/tests.log
python tools/exp_steps_all.py 50000 2>&1 | tail -1 | cut -c1-1200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/c3_rows_launches.csv python tools/exp_c3_steps.py 50000 23 > $out/ncu_c3.log 2>&1; tail -1 $out/ncu_c3.log
