#!/bin/bash
# r02 experiment 5: v5 with integer accumulation in the epilogue (no fp64 under the MMAs)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5 (correctness)"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5c.txt 2>&1; grep -E "error|mismatch|max .err" $O/ozaki_probe_v5c.txt
echo "== drain probe"; timeout 60 tools/micro/ozaki_probe drain 2>&1 | grep -E "int64|1 accumulator" | head -8
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_XP=24" "EGX_OZAKI_V=5 EGX_OZAKI_XP=4" "EGX_OZAKI_V=5 EGX_OZAKI_PERSIST=0"; do
  echo "== $cfg"
  env EGX_OZAKI_PERSIST=1 $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n 7 | grep -E "Mt=|v5 CTA|second tile, [Me]" | tee -a $O/x5.txt
done
echo "== pytest ozaki + parity"; timeout 600 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 -x 2>&1 | tail -5
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5"; do
  echo "== batch sweep 8192: $cfg"
  env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x5_batch.txt
done
