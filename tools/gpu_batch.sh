#!/bin/bash
# scratch: the GPU-box side of one gpurun call (edited per experiment)
out=gpurun_out/$1; mkdir -p $out
python -m pytest tests -x -q -m gpu --durations=6 > $out/gpu_tests.log 2>&1; tail -14 $out/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
