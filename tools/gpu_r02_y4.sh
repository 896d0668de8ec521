#!/bin/bash
# r02 experiment y4: K3 r02 form with 8-deep write-back and 16-byte loads in the inverse; all-L-prefetch K5; batch knobs
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== potrf probe v2"; PROBE_V=2 timeout 120 tools/micro/potrf_probe 2>&1 | grep -E "stamps|us per launch|probe:|max|info" | tail -8 | tee -a $O/y4_potrf_probe.txt
echo "== pytest parity + ozaki + fullsize + fit_api + sgp + moe"; timeout 1200 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_fit_api.py tests/test_gpu_sgp.py tests/test_gpu_moe.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -8
echo "== single eval profile 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -3 | cut -c1-900 | tee -a $O/y4_single.txt
for cfg in "X=0" "EGX_OZAKI_MAXCTAS=132" "EGX_OZAKI_MAXCTAS=148" "EGX_TRSM_ROWS=64"; do
  echo "== batch sweep 8192: $cfg"; env $cfg timeout 300 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | tee -a $O/y4_batch.txt
done
for cfg in "X=0" "EGX_BATCH_LOOKAHEAD=0"; do
echo "== midsize $cfg"; env $cfg timeout 300 python tools/midsize_probe.py 1000 2000 2>&1 | tail -2 | cut -c1-400 | tee -a $O/y4_midsize.txt
echo "== C4-like $cfg"; env $cfg timeout 300 python tools/configs_probe.py c4 2>&1 | tail -1 | tee -a $O/y4_c4.txt
done
echo "== grad probe"; timeout 300 python tools/grad_probe.py 2>&1 | tail -2 | cut -c1-1500 | tee $O/y4_grad.txt
