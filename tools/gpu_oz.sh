#!/bin/bash
# tcgen05 (int8-sliced) trailing update vs the DMMA kernel inside the real sweep: parity tests, then the n = 8192 probe
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (EGX_OZAKI default)"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x 2>&1 | tail -30 > gpurun_out/pytest_oz.log; tail -12 gpurun_out/pytest_oz.log
for cfg in "EGX_OZAKI=0" "EGX_OZAKI=1 EGX_OZAKI_PERSIST=0" "EGX_OZAKI=1 EGX_OZAKI_PERSIST=1" "EGX_OZAKI=1 EGX_OZAKI_2CTA=1 EGX_OZAKI_PERSIST=1"; do
  echo "== probe 8192: $cfg"
  env $cfg timeout 300 python tools/gpu_probe.py 8192 2>&1 | grep -v predict_valvar | cut -c1-700
done
echo "== probe 4096"; EGX_OZAKI=0 timeout 200 python tools/gpu_probe.py 4096 2>&1 | grep -E "batch12|noprof" | cut -c1-300; EGX_OZAKI=1 timeout 200 python tools/gpu_probe.py 4096 2>&1 | grep -E "batch12|noprof" | cut -c1-700
