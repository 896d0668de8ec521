#!/bin/bash
# r02 experiment y14: a fit with 2 / 3 chains at n = 8192 (what a rank of the 8- / 4-GPU e2e leg carries): grid cap of the update
# kernel, look-ahead on / off
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for ns in 1 2; do
for cfg in "X=0" "EGX_OZAKI_MAXCTAS=100" "EGX_OZAKI_MAXCTAS=148" "EGX_BATCH_LOOKAHEAD=0" "EGX_GRAPHS=0"; do
  echo "== n_start=$ns $cfg"; env PROBE_NSTART=$ns $cfg timeout 200 python tools/fit_probe.py 8192 2>&1 | tail -1 | cut -c1-220 | tee -a $O/y14_fit_few_chains.txt
done
done
