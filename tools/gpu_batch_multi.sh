#!/bin/bash
# scratch: multi-GPU side of one gpurun call: $1 = output tag, $2 = number of GPUs
out=gpurun_out/$1; mkdir -p $out; N=$2
if [ "$3" == "tests" ]; then
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $out/multi_tests.log 2>&1; tail -8 $out/multi_tests.log
python -m pytest tests/test_gpu_host_shim.py -x -q -m gpu -k "sharded" > $out/shim_sharded.log 2>&1; tail -8 $out/shim_sharded.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err; tail -c 400 $out/bench_n$N.err
python - $out/bench_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ('n_gpus','value','ms_per_step','gpu_launches')}, d['roofline']['frac'])
    print('e2e', d.get('e2e')); print('parity', d.get('parity'))
except Exception as e: print('bench parse failed', e)
PY
