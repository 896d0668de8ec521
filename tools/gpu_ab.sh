#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
echo "== probe 8192"; timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof|stage_ms" | cut -c1-420
echo "== probe 4096"; timeout 200 python tools/gpu_probe.py 4096 2>&1 | grep -E "batch12|noprof" | cut -c1-300
