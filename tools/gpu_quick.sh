#!/bin/bash
# quick validation session: parity tests + one probe + short bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -45 > gpurun_out/pytest_gpu.log; tail -45 gpurun_out/pytest_gpu.log
echo "== sgp probe"; timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee gpurun_out/sgp_probe.log | tail -4
echo "== probe"; for w in 1 2 3 4; do EGX_BATCH_STREAMS=$w timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof" | cut -c1-300; done; EGX_LOOKAHEAD=0 EGX_BATCH_STREAMS=3 timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof" | cut -c1-300; timeout 200 python tools/gpu_probe.py 2048 8192 2>&1 | tee gpurun_out/probe_quick.log | cut -c1-900
echo "== short bench"; timeout 600 python bench.py --steps 2 --warmup 3 --evals 101 --npred 20000 2>&1 | tee gpurun_out/bench_short.log | cut -c1-3000
