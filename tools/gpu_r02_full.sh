#!/bin/bash
# full GPU validation: every -m gpu test, smoke, short bench (N = 1)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -25 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench short"; timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 2>&1 | tee $O/bench_short.log | cut -c1-6000
