#!/bin/bash
# r02 experiment y6: graph replay above n = 4096 under the look-ahead schedule; warm fit through the public API
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest chain + parity + fullsize + fit_api + theta_grad"; timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fit_api.py tests/test_gpu_theta_grad.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -8
echo "== single eval 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -2 | tail -1 | tee -a $O/y6_single.txt
echo "== fit 8192"; timeout 300 python tools/fit_probe.py 8192 2>&1 | tail -2 | tee -a $O/y6_fit.txt
echo "== fit 8192, cache 12 GB"; EGX_CACHE_MB=12000 timeout 300 python tools/fit_probe.py 8192 2>&1 | tail -2 | tee -a $O/y6_fit.txt
echo "== batch 8192 x 96"; timeout 300 python tools/batch_sweep.py 8192 96 2>&1 | tail -1 | tee -a $O/y6_batch.txt
echo "== grad probe"; timeout 300 python tools/grad_probe.py 2>&1 | tail -2 | cut -c1-400 | tee $O/y6_grad.txt
