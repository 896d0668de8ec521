#!/bin/bash
# r02 experiment 1: what bounds the tcgen05 update kernel -- shared-memory port or the L2 -> SM path?
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
echo "== bulk probe"; timeout 120 tools/micro/bulk_probe > $O/bulk_probe.txt 2>&1; cat $O/bulk_probe.txt
echo "== i8mma probe"; timeout 120 tools/micro/i8mma_probe > $O/i8mma_probe.txt 2>&1; cat $O/i8mma_probe.txt
echo "== ozaki probe (correctness, default)"; timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_default.txt 2>&1; cat $O/ozaki_probe_default.txt
echo "== ozaki probe (correctness, NW0=3)"; EGX_OZAKI_NW0=3 timeout 120 tools/micro/ozaki_probe > $O/ozaki_probe_nw3.txt 2>&1; cat $O/ozaki_probe_nw3.txt
for persist in 0 1; do
for cfg in "EGX_OZAKI_XP=0" "EGX_OZAKI_NW0=3" "EGX_OZAKI_XP=1" "EGX_OZAKI_XP=2" "EGX_OZAKI_XP=4" "EGX_OZAKI_XP=8" "EGX_OZAKI_XP=16" "EGX_OZAKI_XP=24" "EGX_OZAKI_XP=28" "EGX_OZAKI_XP=29" "EGX_OZAKI_XP=25"; do
  echo "== persist=$persist $cfg"
  env EGX_OZAKI_PERSIST=$persist $cfg timeout 60 tools/micro/ozaki_probe time 2>&1 | tee -a $O/xp_persist${persist}.txt
done; done
