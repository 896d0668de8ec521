#!/bin/bash
# r02 experiment y20: C5 knobs after the look-ahead change
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for cfg in "X=0" "EGX_OZAKI_MIN_TRI=2" "EGX_BATCH_STREAMS=12" "EGX_BATCH_STREAMS=24" "EGX_OZAKI_MIN_TRI=2 EGX_BATCH_STREAMS=24"; do
echo "== C5 $cfg"; env $cfg timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y20_c5.txt
done
