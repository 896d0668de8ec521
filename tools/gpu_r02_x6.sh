#!/bin/bash
# r02 experiment 6: v5 (conversion overlapped with the TMA read) in the probe and in the factorisation pipeline
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5 (correctness)"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5d.txt 2>&1; grep -E "error|mismatch|max .err" $O/ozaki_probe_v5d.txt
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_XP=16" "EGX_OZAKI_V=5 EGX_OZAKI_XP=4"; do
  echo "== $cfg"
  env EGX_OZAKI_PERSIST=1 $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n 7 | grep -E "Mt=|v5 CTA|second tile, [Me]" | tee -a $O/x6.txt
done
echo "== pytest ozaki + parity + fullsize"; timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --timeout 400 -x 2>&1 | tail -5
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_BATCH_STREAMS=6" "EGX_OZAKI_V=5 EGX_BATCH_STREAMS=10"; do
  echo "== batch sweep 8192: $cfg"
  env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x6_batch.txt
done
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_PERSIST=1"; do
  echo "== single eval 8192: $cfg"
  env $cfg timeout 300 python tools/gpu_probe.py 8192 2>&1 | grep -E "noprof|predict_valvar" | cut -c1-400 | tee -a $O/x6_single.txt
done
