#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 300 $out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/c3_rows_launches.csv python tools/exp_c3_steps.py 50000 23 > $out/ncu_c3.log 2>&1; tail -1 $out/ncu_c3.log
ncu --set full --clock-control none --import-source on -k regex:row_walk -s 2 -c 1 -o $out/c3_rowwalk_full python tools/exp_c3_steps.py 50000 14 > $out/ncu_c3_full.log 2>&1; tail -1 $out/ncu_c3_full.log
ncu --set full --clock-control none --import-source on -k regex:row_build -s 1 -c 1 -o $out/c3_rowbuild_full python tools/exp_c3_steps.py 50000 22 > $out/ncu_c3_b.log 2>&1; tail -1 $out/ncu_c3_b.log
python tools/exp_steps_all.py 200000 > $out/exp_steps_200k.txt 2>&1; python tools/exp_steps_all.py 50000 > $out/exp_steps_50k.txt 2>&1; tail -n 1 $out/exp_steps_50k.txt; tail -n 1 $out/exp_steps_200k.txt
