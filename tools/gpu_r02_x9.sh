#!/bin/bash
# r02 experiment 9: pre-scaled correlation kernels + sparse GP on tcgen05 -- parity, C3 / C5 probes
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest (all gpu tests)"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -15 | tee $O/pytest_gpu.log
echo "== sgp probe (C3) tcgen05"; timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee $O/x9_sgp_oz.txt | cut -c1-600
echo "== sgp probe (C3) DMMA"; EGX_OZAKI=0 timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee $O/x9_sgp_dmma.txt | cut -c1-600
for w in 12 16 24 32; do
  echo "== C5, W=$w"; EGX_BATCH_STREAMS=$w timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/x9_c5.txt
done
echo "== batch sweep 8192"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/x9_batch.txt
