#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
for lc in 1 0; do echo "== probe 8192 LATEC=$lc"; EGX_GEMM_LATEC=$lc timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof|stage_ms" | cut -c1-420; done
echo "== midsize"; timeout 300 python tools/midsize_probe.py 2>&1 | tee gpurun_out/midsize_graphs.log | cut -c1-400
echo "== sgp probe"; timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee gpurun_out/sgp_probe.log | tail -3 | cut -c1-500
