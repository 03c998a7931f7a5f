#!/bin/bash
# scratch: the GPU-box side of one gpurun call (edited per experiment)
out=gpurun_out/$1; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu > $out/tests.log 2>&1; tail -3 $out/tests.log
python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 300 $out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; cut -c1-200 $out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/c5_launches.csv python bench.py --steps 2 --warmup 3 --no-others --no-cpu-baseline --no-parity --no-e2e > $out/ncu_c5.log 2>&1; tail -1 $out/ncu_c5.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/c3_rows_launches.csv python tools/exp_c3_steps.py 50000 23 > $out/ncu_c3.log 2>&1; tail -1 $out/ncu_c3.log
ncu --set full --clock-control none --import-source on -k regex:row_walk -s 2 -c 1 -o $out/c3_rowwalk_full python tools/exp_c3_steps.py 50000 14 > $out/ncu_c3_full.log 2>&1; tail -1 $out/ncu_c3_full.log
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o $out/tric_full python tools/exp_tric.py 3 > $out/ncu_tric.log 2>&1; tail -1 $out/ncu_tric.log
python tools/exp_cells.py 50000 > $out/exp_cells_50k.txt 2>&1; python tools/exp_steps_all.py 200000 > $out/exp_steps_200k.txt 2>&1; python tools/exp_steps_all.py 50000 > $out/exp_steps_50k.txt 2>&1; tail -1 $out/exp_steps_50k.txt $out/exp_steps_200k.txt
python tools/exp_tric.py 4 | tail -1
