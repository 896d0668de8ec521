#!/bin/bash
# r02 experiment y9: K5 with its own T buffer and unrolled K loop, diagonal-tile update with all loads up front
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest chain + parity + fullsize + sgp + ozaki + fit_api"; timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sgp.py tests/test_gpu_ozaki.py tests/test_gpu_fit_api.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -6
echo "== single eval 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -2 | cut -c1-800 | tee -a $O/y9_single.txt
echo "== batch 8192 x 48"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y9_batch.txt
echo "== batch 8192 x 48, cap 132"; EGX_OZAKI_MAXCTAS=132 timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y9_batch.txt
echo "== C5"; timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y9_c5.txt
echo "== C4"; timeout 300 python tools/configs_probe.py c4 2>&1 | tail -1 | tee -a $O/y9_c4.txt
echo "== sgp"; PROBE_NOPROF=1 timeout 300 python tools/sgp_probe.py 2>&1 | tail -2 | cut -c1-200 | tee -a $O/y9_sgp.txt
echo "== ncu trsm + diag"; timeout 600 ncu --set full --clock-control none -k regex:"trsm_rows_kernel|diag_tile_update" --launch-skip 40 --launch-count 6 -f -o $O/y9_chain python tools/gpu_probe.py 8192 > $O/y9_ncu.log 2>&1; tail -2 $O/y9_ncu.log
