#!/bin/bash
# One GPU-box session: parity tests, A/B probes, ncu captures, short bench.  Outputs -> gpurun_out/
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 240 -s 2>&1 | tail -45 > gpurun_out/pytest_gpu.log; tail -45 gpurun_out/pytest_gpu.log
for cfg in "1 64" "0 64" "1 128" "0 128"; do set -- $cfg
  echo "== probe lookahead=$1 gemm_bn=$2"; EGX_LOOKAHEAD=$1 EGX_GEMM_BN=$2 timeout 200 python tools/gpu_probe.py 8192 2>&1 | tee gpurun_out/probe_la$1_bn$2.log | cut -c1-900
done
echo "== probe small sizes"; timeout 200 python tools/gpu_probe.py 256 1024 2048 4096 2>&1 | tee gpurun_out/probe_sizes.log | cut -c1-700
if [ "$1" != "noncu" ]; then
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 8192 2048 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_sub|potrf_diag|trsm_rows|corr_build|cross_corr|gls_kernel|var_finish" -c 14 -f -o gpurun_out/prof_r01 python tools/ncu_target.py 8192 2048 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
fi
echo "== short bench"; timeout 600 python bench.py --steps 2 --warmup 3 --evals 101 --npred 20000 2>&1 | tee gpurun_out/bench_short.log | cut -c1-3000
