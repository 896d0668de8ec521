#!/bin/bash
# r02 experiment y22: the factorisation leaves the slices of L where the solves read them
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest theta_grad + fullsize + parity + fit_api + chain + ozaki"; timeout 1200 python -m pytest tests/test_gpu_theta_grad.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_fit_api.py tests/test_gpu_chain.py tests/test_gpu_ozaki.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4
echo "== grad probe"; timeout 300 python tools/grad_probe.py 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d[k] for k in ('n', 'd', 'closed_form_ms', 'central_differences_ms', 'max_rel_diff', 'closed_form_status')}))" | tee -a $O/y22_grad.txt
echo "== single eval + predict 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | sed -n 2,5p | cut -c1-300 | tee $O/y22_single.txt
