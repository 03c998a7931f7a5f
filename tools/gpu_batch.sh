#!/bin/bash
# scratch: the GPU-box side of one gpurun call (edited per experiment)
out=gpurun_out/$1; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > $out/tests.log 2>&1; tail -3 $out/tests.log
