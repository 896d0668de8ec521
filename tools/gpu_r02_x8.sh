#!/bin/bash
# r02 experiment 8: final v5 (3-stage ring), pre-scaled correlation kernels, resident-grid cap, batch width at n = 2048
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5f.txt 2>&1; grep -E "error|mismatch|max .err" $O/ozaki_probe_v5f.txt
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5"; do
  echo "== $cfg"; env EGX_OZAKI_PERSIST=1 $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n 7 | grep -E "Mt=|v5 CTA" | tee -a $O/x8.txt
done
echo "== pytest parity + ozaki + fullsize + fit_api + moe"; timeout 1200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fit_api.py tests/test_gpu_moe.py tests/test_gpu_sgp.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -15
for cfg in "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_MAXCTAS=132" "EGX_OZAKI_V=5 EGX_OZAKI_MAXCTAS=120" "EGX_OZAKI_V=5 EGX_OZAKI_MAXCTAS=104" "EGX_OZAKI_V=5 EGX_OZAKI_MAXCTAS=120 EGX_BATCH_STREAMS=10"; do
  echo "== batch sweep 8192: $cfg"
  env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x8_batch.txt
done
echo "== single eval profile 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -3 | cut -c1-700 | tee $O/x8_single.txt
for w in 12 16 24 32; do
  echo "== C5 n=2048, W=$w"; EGX_BATCH_STREAMS=$w timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/x8_c5.txt
done
