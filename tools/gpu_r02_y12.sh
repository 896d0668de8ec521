#!/bin/bash
# r02 experiment y12: K1 with one CTA per 128 x 128 block
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest parity + chain + sgp + moe"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py tests/test_gpu_sgp.py tests/test_gpu_moe.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4
echo "== corr probe, 128-block CTAs"; timeout 300 python tools/corr_probe.py 2>&1 | tee $O/y12_corr_128.txt
echo "== corr probe, 64-tile CTAs"; EGX_CORR_TILE=64 timeout 300 python tools/corr_probe.py 2>&1 | tee $O/y12_corr_64.txt
echo "== batch 8192 x 48"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y12_batch.txt
echo "== C5"; timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y12_c5.txt
