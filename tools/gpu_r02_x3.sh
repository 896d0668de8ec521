#!/bin/bash
# r02 experiment 3: the four-group kernel (v5) -- correctness, timing, inside the factorisation
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5 (correctness)"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5.txt 2>&1; grep -E "error|mismatch|max .err|Mt=|v5|second tile" $O/ozaki_probe_v5.txt
for persist in 1 0; do
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=4" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_XP=4" "EGX_OZAKI_V=5 EGX_OZAKI_XP=24" "EGX_OZAKI_V=5 EGX_OZAKI_XP=28" "EGX_OZAKI_V=5 EGX_OZAKI_XP=1"; do
  echo "== persist=$persist $cfg"
  env EGX_OZAKI_PERSIST=$persist $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n +6 | tee -a $O/x3_persist${persist}.txt
done; done
echo "== pytest ozaki + parity"; timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 -x 2>&1 | tail -5
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5"; do
  echo "== probe 8192: $cfg"
  env $cfg timeout 300 python tools/gpu_probe.py 8192 2>&1 | grep -v predict_valvar | cut -c1-900 | tee -a $O/x3_probe8192.txt
done
