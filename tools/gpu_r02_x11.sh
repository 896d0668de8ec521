#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe v5"; EGX_OZAKI_V=5 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_v5g.txt 2>&1; grep -E "error|mismatch|max .err" $O/ozaki_probe_v5g.txt
for cfg in "EGX_OZAKI_V=3" "EGX_OZAKI_V=5" "EGX_OZAKI_V=5 EGX_OZAKI_XP=16"; do
  echo "== $cfg"; env EGX_OZAKI_PERSIST=1 $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tail -n 7 | grep -E "Mt=|v5 CTA|second tile, [Me]" | tee -a $O/x11.txt
done
echo "== pytest parity + ozaki + fullsize"; timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sgp.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4
echo "== batch sweep 8192"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x11_batch.txt
echo "== C5"; timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/x11_c5.txt
