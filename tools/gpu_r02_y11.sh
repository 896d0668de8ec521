#!/bin/bash
# r02 experiment y11: racecheck after the K3 barrier fix, sparse tests, K3 probe
mkdir -p gpurun_out/r02
O=gpurun_out/r02
export PYTHONUNBUFFERED=1
echo "== potrf probe"; PROBE_V=2 timeout 120 tools/micro/potrf_probe 2>&1 | grep -E "stamps|us per launch|probe:|max" | tail -4 | tee $O/y11_potrf_probe.txt
echo "== pytest sgp + chain"; timeout 900 python -m pytest tests/test_gpu_sgp.py tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider --timeout 400 2>&1 | tail -4
echo "== compute-sanitizer racecheck"; EGX_GEMM_MB=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
