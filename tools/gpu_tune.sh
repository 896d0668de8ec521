#!/bin/bash
export PYTHONUNBUFFERED=1
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_ozaki.py -q -p no:cacheprovider --timeout 300 2>&1 | tail -15
for n in 2048 2560 3072 3584; do for oz in 0 1; do
  echo "== n=$n EGX_OZAKI=$oz (MIN_T=1)"; EGX_OZAKI_MIN_T=1 EGX_OZAKI=$oz timeout 200 python tools/batch_sweep.py $n 96 2>&1 | tail -1 | cut -c1-200
done; done
