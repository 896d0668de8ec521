#!/bin/bash
# bench on N GPUs of one box (strong scaling of the headline step + the c5 / c4 legs), launched the way the driver launches it
N=${1:-4}
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export PYTHONUNBUFFERED=1
nvidia-smi -L
if [ "$N" = "2" ]; then echo "== pytest parallel (2 GPUs: NCCL variant runs)"; timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4; fi
echo "== bench N=$N"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 3 > $O/bench_${N}gpu.log 2>$O/bench_${N}gpu.err; tail -3 $O/bench_${N}gpu.err; python - $N <<'PY'
import json, sys
n = sys.argv[1]
for line in open('gpurun_out/r02/bench_%sgpu.log' % n):
    if line.startswith('{'):
        o=json.loads(line)
        print({k:o.get(k) for k in ('value','n_gpus','ms_per_step','device_ms_per_step','likelihood_evals_per_s','scaling','sweep_allgather_us_rank0')})
        print('e2e',o['e2e']); print('c5',o.get('c5_theta_sweep')); print('c4',o.get('c4_moe_experts'))
PY
echo "== reference arm N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-600
