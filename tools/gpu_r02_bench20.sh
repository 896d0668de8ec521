#!/bin/bash
# the bench as the driver runs it (20 steps), with the wall time of the command
mkdir -p gpurun_out/r02
SECONDS=0
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02/bench_20steps.log 2> gpurun_out/r02/bench_20steps.err
echo "rc=$? wall_s=$SECONDS"
tail -1 gpurun_out/r02/bench_20steps.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'steps', 'clocks', 'gpu_launches')})
print(d['e2e'])"
tail -2 gpurun_out/r02/bench_20steps.err
