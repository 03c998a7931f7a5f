#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
python tools/exp_steps_all.py 50000 2>&1 | tail -2
python tools/exp_steps_all.py 200000 2>&1 | tail -2
