#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest (all gpu tests)"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -12 | tee $O/pytest_gpu.log
echo "== sgp probe (C3) tcgen05"; timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee $O/x10_sgp_oz.txt | cut -c1-600
echo "== bench short"; timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 2>&1 | tee $O/bench_short.log | cut -c1-200; python - <<'PY'
import json
for line in open('gpurun_out/r02/bench_short.log'):
    if line.startswith('{'):
        o=json.loads(line)
        print({k:o[k] for k in ('value','ms_per_step','device_ms_per_step','likelihood_evals_per_s')})
        print('e2e',o['e2e']); print('c3',o.get('c3_sparse_gp')); print('c5',o.get('c5_theta_sweep')); print('c4',o.get('c4_moe_experts'))
        r=o['roofline']; print({k:r[k] for k in ('achieved','tensor_pipe_frac','avg_launch_ms','launches','executed_int8_tops')}); print(r.get('corr_build_kernel')); print(o.get('cpu_baseline'))
PY
