#!/bin/bash
# r02 experiment y7: e2e breakdown in the bench line, sparse-GP chunk size
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for ch in 8192 9472 18944; do
echo "== sgp chunk $ch"; PROBE_NOPROF=1 EGX_SGP_CHUNK=$ch timeout 300 python tools/sgp_probe.py 2>&1 | tail -2 | cut -c1-300 | tee -a $O/y7_sgp.txt
done
echo "== pytest sgp"; EGX_SGP_CHUNK=9472 timeout 600 python -m pytest tests/test_gpu_sgp.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -3
echo "== bench short"; timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-extra 2>&1 | tail -1 > $O/y7_bench_short.log; python - <<'P'
import json
d=json.loads(open("gpurun_out/r02/y7_bench_short.log").read())
print({k:d[k] for k in ("value","ms_per_step","e2e","fit_ms_per_step_rank0","predict_ms_per_step_rank0","clocks") if k in d})
P
