#!/bin/bash
# r02 experiment y18: look-ahead column updates on tcgen05 (slices shared with the bulk update, double-buffered)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest chain + parity + fullsize + ozaki + fit_api + moe + theta_grad"; timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_ozaki.py tests/test_gpu_fit_api.py tests/test_gpu_moe.py tests/test_gpu_theta_grad.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -5
for cfg in "X=0" "EGX_LA_OZAKI=0"; do
echo "== single eval 8192 $cfg"; env $cfg timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -2 | tail -1 | tee -a $O/y18_single.txt
echo "== C4 $cfg"; env $cfg timeout 300 python tools/configs_probe.py c4 2>&1 | tail -1 | cut -c1-200 | tee -a $O/y18_c4.txt
echo "== fit 2 chains $cfg"; env PROBE_NSTART=1 $cfg timeout 200 python tools/fit_probe.py 8192 2>&1 | tail -1 | cut -c1-130 | tee -a $O/y18_fit2.txt
echo "== batch 4096 x 96 $cfg"; env $cfg timeout 300 python tools/batch_sweep.py 4096 96 2>&1 | tail -1 | tee -a $O/y18_batch4096.txt
done
