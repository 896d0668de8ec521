#!/bin/bash
# r02 experiment y10: sparse-GP trajectories, shared sampler; batch width / look-ahead at n = 8192 after the chain work
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (all)"; timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -12 | tee $O/y10_pytest_gpu.log
for cfg in "X=0" "EGX_BATCH_STREAMS=6" "EGX_BATCH_STREAMS=10" "EGX_BATCH_LOOKAHEAD=1 EGX_GRAPHS=0" "EGX_BATCH_LOOKAHEAD=1"; do
  echo "== batch sweep 8192: $cfg"; env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y10_batch.txt
done
