#!/bin/bash
# r02 experiment y16: staggered chain starts with 2 / 3 chains at n = 8192
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for us in 0 1200 2400 3600; do
  echo "== n_start=1 stagger $us"; env PROBE_NSTART=1 EGX_FIT_STAGGER_US=$us timeout 200 python tools/fit_probe.py 8192 2>&1 | tail -1 | cut -c1-130 | tee -a $O/y16_stagger.txt
done
for us in 0 1000 1600 2400; do
  echo "== n_start=2 stagger $us"; env PROBE_NSTART=2 EGX_FIT_STAGGER_US=$us timeout 200 python tools/fit_probe.py 8192 2>&1 | tail -1 | cut -c1-130 | tee -a $O/y16_stagger.txt
done
