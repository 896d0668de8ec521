#!/bin/bash
# r02 experiment y2: phase clock stamps of the r02 K3, its variants; diagonal-tile-first look-ahead schedule
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for b in potrf_probe potrf_probe_one_inst potrf_probe_librsqrt; do echo "== $b"; timeout 120 tools/micro/$b 2>&1 | grep -E "stamps|us per launch|probe:|max" | tail -4 | tee -a $O/y2_potrf_probe.txt; done
echo "== pytest parity + ozaki + fullsize + fit_api"; timeout 1200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fit_api.py -m gpu -q -x -p no:cacheprovider --timeout 400 2>&1 | tail -8
for v in 1 2; do
echo "== single eval profile 8192, look-ahead schedule $v"; EGX_LOOKAHEAD_V=$v timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -3 | cut -c1-900 | tee -a $O/y2_single.txt
echo "== C5, look-ahead schedule $v"; EGX_LOOKAHEAD_V=$v timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y2_c5.txt
done
echo "== grad probe"; timeout 300 python tools/grad_probe.py 2>&1 | tail -2 | cut -c1-1500 | tee $O/y2_grad.txt
echo "== midsize"; timeout 300 python tools/midsize_probe.py 2>&1 | tail -6 | tee $O/y2_midsize.txt
