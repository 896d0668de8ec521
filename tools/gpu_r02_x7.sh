#!/bin/bash
# r02 experiment 7: v5 with the byte-granular operand ring; ncu evidence
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5 (correctness)"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5e.txt 2>&1; grep -E "error|mismatch|max .err" $O/ozaki_probe_v5e.txt
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_XP=16" "EGX_OZAKI_V=5 EGX_OZAKI_XP=4"; do
  echo "== $cfg"
  env EGX_OZAKI_PERSIST=1 $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n 7 | grep -E "Mt=|v5 CTA|second tile, [Me]" | tee -a $O/x7.txt
done
echo "== pytest ozaki + parity + fullsize"; timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --timeout 400 -x 2>&1 | tail -4
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5"; do
  echo "== batch sweep 8192: $cfg"
  env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x7_batch.txt
done
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 8192 2048 > $O/ncu_list.log 2>&1; tail -2 $O/ncu_list.log
echo "== ncu full (ozaki_syrk5)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ozaki_syrk5" -s 2 -c 3 -f -o gpurun_out/prof_r02 python tools/ncu_target.py 8192 2048 > $O/ncu_full.log 2>&1; tail -3 $O/ncu_full.log
ls -la gpurun_out/*.ncu-rep
