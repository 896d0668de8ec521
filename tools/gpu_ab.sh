#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
echo "== rounds cache"; timeout 300 python tools/round_probe.py 300 500 1000 2000 2>&1 | cut -c1-500
echo "== rounds no cache"; EGX_CACHE_MB=0 timeout 300 python tools/round_probe.py 300 500 1000 2000 2>&1 | cut -c1-500
