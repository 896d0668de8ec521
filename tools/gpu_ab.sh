#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== smoke MB"; EGX_GEMM_MB=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== tests MB"; EGX_GEMM_MB=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sgp.py -m gpu -q -x -p no:cacheprovider --timeout 200 2>&1 | tail -5
for mb in 0 1 0 1; do EGX_GEMM_MB=$mb timeout 200 python tools/batch_sweep.py 8192 48 2>&1 | tail -1 | sed "s/^/MB=$mb /"; done
for mb in 0 1; do echo "== probe MB=$mb"; EGX_GEMM_MB=$mb timeout 200 python tools/gpu_probe.py 8192 2>&1 | grep -E "noprof|stage_ms" | cut -c1-330; done
