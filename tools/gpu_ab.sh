#!/bin/bash
export PYTHONUNBUFFERED=1
for w in 2 3 4 5 6 8; do EGX_BATCH_STREAMS=$w timeout 200 python tools/batch_sweep.py 8192 48 2>&1 | tail -1; done
for w in 4 6 8 12; do EGX_BATCH_STREAMS=$w timeout 200 python tools/batch_sweep.py 4096 96 2>&1 | tail -1; done
