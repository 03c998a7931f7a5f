#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $out/multi_tests.log 2>&1; tail -4 $out/multi_tests.log
timeout 900 python -m pytest tests/test_gpu_host_shim.py -q -m gpu -k "sharded" > $out/shim_sharded.log 2>&1; tail -6 $out/shim_sharded.log
