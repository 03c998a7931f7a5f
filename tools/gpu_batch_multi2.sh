#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $out/multi_tests.log 2>&1; tail -4 $out/multi_tests.log
timeout 900 python -m pytest tests/test_gpu_host_shim.py -q -m gpu -k "sharded" > $out/shim_sharded.log 2>&1; tail -6 $out/shim_sharded.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err; tail -c 400 $out/bench_n2.json
