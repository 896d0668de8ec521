#!/bin/bash
# r02 experiment y5: new chain tests, CUDA-graph replay at n = 8192, fit through the public API, sparse GP, short bench
mkdir -p gpurun_out/r02
O=gpurun_out/r02
echo "== pytest chain"; timeout 900 python -m pytest tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -8
for cfg in "X=0" "EGX_GRAPHS=1"; do
echo "== single eval 8192 $cfg"; env $cfg timeout 300 python tools/gpu_probe.py 8192 2>&1 | head -3 | tail -2 | tee -a $O/y5_single.txt
echo "== fit 8192 $cfg"; env $cfg timeout 300 python tools/fit_probe.py 8192 2>&1 | tail -1 | tee -a $O/y5_fit.txt
done
echo "== C5"; timeout 300 python tools/configs_probe.py c5 2>&1 | tail -1 | tee -a $O/y5_c5.txt
echo "== sgp"; timeout 300 python tools/sgp_probe.py 2>&1 | tail -2 | cut -c1-900 | tee $O/y5_sgp.txt
echo "== bench short"; timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-extra 2>&1 | tail -1 | cut -c1-3000 | tee $O/y5_bench_short.log
