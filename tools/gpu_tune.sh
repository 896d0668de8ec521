#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fit_api.py tests/test_gpu_ozaki.py tests/test_gpu_fullsize.py -q -p no:cacheprovider --timeout 300 2>&1 | tail -6
for oz in 0 1; do echo "== probe 8192 EGX_OZAKI=$oz"; EGX_OZAKI=$oz timeout 300 python tools/gpu_probe.py 8192 2>&1 | grep predict_valvar | cut -c1-700; done
