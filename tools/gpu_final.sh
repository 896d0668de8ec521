#!/bin/bash
# Full evidence session on one B200: parity tests, compute-sanitizer, secondary configs, ncu captures, default bench.
TAG=${1:-r01e}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== compute-sanitizer memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_memcheck.log
# racecheck models bar.sync, not the mbarrier-ordered cp.async pipeline of the DMMA kernel (EGX_GEMM_MB=1, default): it is run on
# the bar.sync variant of that kernel (same arithmetic, same results); the tcgen05 / TMA kernels run as shipped
echo "== compute-sanitizer racecheck"; EGX_GEMM_MB=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/sanitizer_racecheck.log
echo "== configs"; timeout 600 python tools/configs_probe.py c1 c5 c4 2>&1 | tee gpurun_out/configs_probe.log | cut -c1-400
echo "== sgp probe"; timeout 300 python tools/sgp_probe.py 100000 6 1024 2>&1 | tee gpurun_out/sgp_probe.log | tail -3 | cut -c1-500
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 8192 2048 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ozaki_syrk|ozaki_slice|ozaki_rowscale|gemm_nt_sub|potrf_diag|trsm_rows|corr_build|cross_corr|gls_kernel|var_finish|backsolve|diag_tile_update" -c 30 -f -o gpurun_out/prof_$TAG python tools/ncu_target.py 8192 2048 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
echo "== bench (default)"; timeout 1200 python bench.py --steps ${STEPS:-5} --warmup 3 2>&1 | tee gpurun_out/bench_default.log | cut -c1-4000
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tee gpurun_out/bench_reference.log | cut -c1-1500
