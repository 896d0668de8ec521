#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export PYTHONUNBUFFERED=1
nvidia-smi -L
echo "== pytest parallel (2 GPUs: NCCL variant runs)"; timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -6
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 2 > $O/bench_2gpu.log 2>$O/bench_2gpu.err; tail -3 $O/bench_2gpu.err; python - <<'PY'
import json
for line in open('gpurun_out/r02/bench_2gpu.log'):
    if line.startswith('{'):
        o=json.loads(line)
        print({k:o[k] for k in ('value','n_gpus','ms_per_step','device_ms_per_step','likelihood_evals_per_s','scaling')})
        print('e2e',o['e2e']); print('c5',o.get('c5_theta_sweep')); print('c4',o.get('c4_moe_experts'))
PY
