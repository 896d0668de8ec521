#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
echo "== probe 8192"; timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof|stage_ms" | cut -c1-420
echo "== sgp probe"; timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee gpurun_out/sgp_probe.log | tail -3 | cut -c1-500
echo "== midsize"; timeout 300 python tools/round_probe.py 500 1000 2>&1 | cut -c1-400
