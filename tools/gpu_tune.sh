#!/bin/bash
export PYTHONUNBUFFERED=1
for i in 1 2; do timeout 200 python tools/batch_sweep.py 8192 48 2>&1 | tail -1; done
timeout 200 python tools/batch_sweep.py 4096 96 2>&1 | tail -1
