#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_ozaki.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -p no:cacheprovider --timeout 300 2>&1 | tail -8
echo "== batch 8192"; timeout 200 python tools/batch_sweep.py 8192 48 2>&1 | tail -1
echo "== probe 8192"; timeout 300 python tools/gpu_probe.py 8192 2>&1 | grep -v predict_valvar | cut -c1-900
echo "== sanitize target (plain)"; timeout 300 python tools/sanitize_target.py 2>&1 | tail -6
