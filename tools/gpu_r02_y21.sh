#!/bin/bash
# r02 experiment y21: closed-form gradient with W W^T pipelined under the sweep
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest theta_grad + fit_api"; timeout 900 python -m pytest tests/test_gpu_theta_grad.py tests/test_gpu_fit_api.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4
for cfg in "X=0" "EGX_GRAD_PIPELINE=0"; do
echo "== grad probe $cfg"; env $cfg timeout 300 python tools/grad_probe.py 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d[k] for k in ('n', 'd', 'closed_form_ms', 'central_differences_ms', 'max_rel_diff', 'closed_form_status')}))" | tee -a $O/y21_grad.txt
done
echo "== lbfgs probe"; timeout 300 python tools/lbfgs_probe.py 2>&1 | tail -2 | cut -c1-400 | tee $O/y21_lbfgs.txt
