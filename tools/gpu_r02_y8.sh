#!/bin/bash
# r02 experiment y8: two-wave chunks (sparse GP, predict_var), 16 GB block cache; full GPU test suite
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -12 | tee $O/y8_pytest_gpu.log
echo "== sgp"; PROBE_NOPROF=1 timeout 300 python tools/sgp_probe.py 2>&1 | tail -2 | cut -c1-200 | tee -a $O/y8_sgp.txt
echo "== bench short"; timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 3 2>&1 | tail -1 > $O/y8_bench_short.log; python - <<'P'
import json
d=json.loads(open("gpurun_out/r02/y8_bench_short.log").read())
print({k:d[k] for k in ("value","ms_per_step","e2e","fit_ms_per_step_rank0","predict_ms_per_step_rank0","clocks","c3_sparse_gp","c5_theta_sweep","c4_moe_experts") if k in d})
P
