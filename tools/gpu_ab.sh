#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -30 gpurun_out/pytest_gpu.log
echo "== probe 8192"; timeout 200 python tools/gpu_probe.py 8192 2>&1 | tee gpurun_out/probe_ab.log | cut -c1-600
echo "== probe 8192 lookahead off"; EGX_LOOKAHEAD=0 timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "batch12|noprof" | cut -c1-300
echo "== probe 2048/4096"; timeout 200 python tools/gpu_probe.py 2048 4096 2>&1 | grep -E "batch12|noprof|predict_valvar" | cut -c1-300
echo "== short bench"; timeout 600 python bench.py --steps 2 --warmup 3 --evals 101 --npred 20000 2>&1 | tee gpurun_out/bench_short.log | cut -c1-3000
