#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "longer" > $out/tests.log 2>&1; tail -15 $out/tests.log
