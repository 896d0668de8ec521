#!/bin/bash
# r02 experiment y15: few chains in flight at n = 8192 with one tile per CTA in the update kernel (the block scheduler can then slip
# the other evaluation's high-priority chain kernels in between tiles)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for ns in 1 2 0; do
for cfg in "X=0" "EGX_OZAKI_PERSIST=0"; do
  echo "== n_start=$ns $cfg"; env PROBE_NSTART=$ns $cfg timeout 200 python tools/fit_probe.py 8192 2>&1 | tail -1 | cut -c1-200 | tee -a $O/y15_fit_few_chains.txt
done
done
