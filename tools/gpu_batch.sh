#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "general or tric or mixed or orthogonal or cell" > $out/tests.log 2>&1; tail -3 $out/tests.log
for v in "" _gu1 _gu4 _gu8; do
  echo "== variant '$v'"
  B200CV_LIB=colvars_b200/libb200cv$v.so python tools/exp_tric.py 4 2>&1 | tail -2
done
