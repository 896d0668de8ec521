#!/bin/bash
# r02 experiment y3: compact-code K3 (third form), all-L-prefetch K5 with 32-row slabs, C5 knobs, ncu of the chain kernels + slicing
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for v in 3 2; do echo "== potrf probe v$v"; PROBE_V=$v timeout 120 tools/micro/potrf_probe 2>&1 | grep -E "stamps|us per launch|probe:|max|info" | tail -8 | tee -a $O/y3_potrf_probe.txt; done
echo "== pytest parity + ozaki + fullsize + fit_api + sgp"; timeout 1200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fit_api.py tests/test_gpu_sgp.py -m gpu -q -x -p no:cacheprovider --timeout 400 2>&1 | tail -8
echo "== single eval profile 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -3 | cut -c1-900 | tee -a $O/y3_single.txt
echo "== single eval 8192, 64-row slabs"; EGX_TRSM_ROWS=64 timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -2 | tail -1 | tee -a $O/y3_single.txt
for cfg in "X=0" "EGX_BATCH_LOOKAHEAD=0" "EGX_OZAKI_MIN_TRI=4" "EGX_BATCH_LOOKAHEAD=0 EGX_OZAKI_MIN_TRI=4" "EGX_BATCH_LOOKAHEAD=0 EGX_OZAKI_MIN_TRI=2"; do
echo "== C5 $cfg"; env $cfg timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y3_c5.txt
done
echo "== grad probe"; timeout 300 python tools/grad_probe.py 2>&1 | tail -2 | cut -c1-1500 | tee $O/y3_grad.txt
echo "== ncu chain kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ozaki_slice_kernel|trsm_rows_kernel|potrf_diag3|diag_tile_update" --launch-skip 60 --launch-count 10 -f -o $O/y3_chain python tools/gpu_probe.py 8192 > $O/y3_ncu.log 2>&1; tail -3 $O/y3_ncu.log
