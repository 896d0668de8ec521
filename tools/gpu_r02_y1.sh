#!/bin/bash
# r02 experiment y1: barrier-free diagonal-block factorisation (K3 r02 form) and the one-launch back substitution
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for v in 1 2; do echo "== potrf probe v$v"; PROBE_V=$v timeout 120 tools/micro/potrf_probe 2>&1 | tee -a $O/y1_potrf_probe.txt; done
echo "== pytest parity + ozaki + fullsize + fit_api + sgp"; timeout 1200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fit_api.py tests/test_gpu_sgp.py -m gpu -q -x -p no:cacheprovider --timeout 400 2>&1 | tail -15
echo "== single eval profile 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -3 | cut -c1-900 | tee $O/y1_single.txt
echo "== grad probe"; timeout 300 python tools/grad_probe.py 2>&1 | tail -2 | cut -c1-1500 | tee $O/y1_grad.txt
echo "== C5"; timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y1_c5.txt
echo "== batch sweep 8192"; timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y1_batch.txt
