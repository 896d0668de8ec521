#!/bin/bash
# r02 experiment y17: slicing kernel with its loads staged through shared memory
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== ozaki probe (accuracy + slice / update times)"; EGX_OZAKI_PERSIST=1 timeout 120 tools/micro/ozaki_probe time 2>&1 | grep -E "Mt=|error|mismatch|max" | head -6 | tee $O/y17_ozaki_probe.txt
echo "== pytest ozaki + parity + fullsize + sgp"; timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sgp.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4
echo "== batch 8192 x 48"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y17_batch.txt
echo "== batch 8192 x 48 again"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y17_batch.txt
echo "== C5"; timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y17_c5.txt
