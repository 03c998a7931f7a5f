#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu > $out/tests.log 2>&1; tail -3 $out/tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 300 $out/bench_n1.err
python - <<PY
import json
d=json.loads(open('$out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['other_workloads']['c3'])
PY
python tools/exp_steps_all.py 200000 | tail -1
