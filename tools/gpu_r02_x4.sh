#!/bin/bash
# r02 experiment 4: v5 with the C tile read at tile start + 32-column TMEM loads
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5 (correctness)"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5b.txt 2>&1; grep -E "error|mismatch|max .err" $O/ozaki_probe_v5b.txt
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=4" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_XP=24" "EGX_OZAKI_V=5 EGX_OZAKI_XP=16" "EGX_OZAKI_V=5 EGX_OZAKI_XP=8" "EGX_OZAKI_V=5 EGX_OZAKI_PERSIST=0"; do
  echo "== $cfg"
  env EGX_OZAKI_PERSIST=1 $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n 7 | grep -E "Mt=|v5|second tile" | tee -a $O/x4.txt
done
echo "== pytest ozaki + parity"; timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 -x 2>&1 | tail -5
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5"; do
  echo "== batch sweep 8192: $cfg"
  env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x4_batch.txt
done
