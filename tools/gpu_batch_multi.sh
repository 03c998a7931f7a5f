#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
N=${2:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err; tail -c 300 $out/bench_n$N.err; python - <<PY
import json
d=json.loads(open('$out/bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['max_rel'], d['roofline']['frac'])
PY
