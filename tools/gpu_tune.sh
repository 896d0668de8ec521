#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_fit_api.py tests/test_gpu_moe.py tests/test_gpu_ozaki.py tests/test_gpu_fullsize.py -q -p no:cacheprovider --timeout 300 2>&1 | tail -6
for ls in 1 0; do echo "== fit probe lockstep=$ls"; EGX_FIT_LOCKSTEP=$ls timeout 300 python tools/fit_probe.py 2000 8192 2>&1 | tail -2; done
