#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
for v in "" _xb296 _xb1184; do
 echo "== $v"; B200CV_LIB=colvars_b200/libb200cv$v.so python tools/exp_steps_all.py 50000 2>&1 | tail -1
done
