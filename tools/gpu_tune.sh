#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -q -p no:cacheprovider --timeout 300 2>&1 | tail -4
for i in 1 2; do timeout 200 python tools/batch_sweep.py 8192 48 2>&1 | tail -1; done
timeout 300 python tools/gpu_probe.py 8192 2>&1 | grep -E "noprof|predict_valvar" | cut -c1-330
