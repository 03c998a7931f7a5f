#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
B200CV_FUZZ_SPARSE=60 timeout 1700 python -m pytest tests/test_gpu_fuzz.py -q -m gpu -k sparse -x > $out/fuzz.log 2>&1; tail -15 $out/fuzz.log
