#!/bin/bash
# scratch: the GPU-box side of one gpurun call (edited per experiment)
out=gpurun_out/$1; mkdir -p $out
for v in "" _m6u4 _m8u2; do echo "== variant $v"; B200CV_LIB=$PWD/colvars_b200/libb200cv$v.so python tools/exp_cells.py 50000 2>&1 | tail -3; done > $out/variants.txt 2>&1; cat $out/variants.txt
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "sparse or overflow or pairlist or fused or cell_list or nearly or periodic or c3" > $out/parity.log 2>&1; tail -4 $out/parity.log
python bench.py --no-e2e --no-cpu-baseline --no-parity --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.err; python - $out <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]+'/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])
    for k,v in d.get('other_workloads',{}).items(): print(k, {a:v.get(a) for a in ('value','ms_per_step','calc_value_ms','graph_step_ms','roofline_frac','pairlist','error')})
except Exception as e: print('bench parse failed', e)
PY
