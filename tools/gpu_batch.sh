#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
for v in "" _nored; do
  echo "== variant '$v'"
  B200CV_LIB=colvars_b200/libb200cv$v.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:row_walk -c 3 --csv --log-file $out/walk$v.csv python tools/exp_c3_steps.py 50000 14 > /dev/null 2>&1
  grep row_walk $out/walk$v.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tail -2
done
