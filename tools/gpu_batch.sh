#!/bin/bash
# scratch: the GPU-box side of one gpurun call (edited per experiment)
out=gpurun_out/$1; mkdir -p $out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sparse or overflow or pairlist or fused or cell_list" > $out/parity.log 2>&1; tail -5 $out/parity.log
python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k c3 > $out/c3.log 2>&1; tail -3 $out/c3.log
python tools/exp_cells.py 50000 > $out/forms.txt 2>&1; cat $out/forms.txt
ncu --set full --clock-control none --import-source on -k regex:row_walk -s 2 -c 1 -o $out/rowwalk_full python tools/exp_c3_steps.py 50000 14 > $out/ncu_full.log 2>&1; tail -2 $out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:row_build -c 1 -o $out/rowbuild_full python tools/exp_c3_steps.py 50000 12 > $out/ncu_full2.log 2>&1; tail -2 $out/ncu_full2.log
