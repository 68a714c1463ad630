#!/bin/bash
TAG=${1:-r2k}
mkdir -p gpurun_out/${TAG}_drivers
DRIVER_HARNESS_KEEP=$PWD/gpurun_out/${TAG}_drivers timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -12 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/grad_diag.py > gpurun_out/${TAG}_grad_diag.log 2>&1; grep -A6 "auto" gpurun_out/${TAG}_grad_diag.log | head -40
IMPLS=bf16,tc timeout 200 python tools/quick.py 2 2>&1 | grep -v Warn | tail -8
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
