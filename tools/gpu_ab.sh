#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
for bk in 16 32; do echo "== GEMM BK=$bk"; EGX_GEMM_BK=$bk timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof|predict_valvar" | cut -c1-420; done
echo "== lookahead off, batch 4"; EGX_LOOKAHEAD=0 timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof" | cut -c1-300
