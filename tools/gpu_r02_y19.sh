#!/bin/bash
# r02 experiment y19: does the look-ahead schedule now pay inside batches (its column updates are on tcgen05)?
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for cfg in "X=0" "EGX_BATCH_LOOKAHEAD=1"; do
echo "== C5 $cfg"; env $cfg timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y19_c5.txt
echo "== batch 8192 x 48 $cfg"; env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y19_batch.txt
echo "== midsize $cfg"; env $cfg timeout 300 python tools/midsize_probe.py 1000 2000 2>&1 | tail -2 | cut -c1-200 | tee -a $O/y19_midsize.txt
done
echo "== batch 4096 x 96 la off"; EGX_BATCH_LOOKAHEAD=0 timeout 300 python tools/batch_sweep.py 4096 96 2>&1 | tail -1 | tee -a $O/y19_batch.txt
