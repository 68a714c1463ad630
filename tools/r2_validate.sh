#!/bin/bash
# Round-2 validation pass on one B200: GPU test suite, bench (both arms), optional extras.
TAG=${1:-r2}
mkdir -p gpurun_out/${TAG}_drivers
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/${TAG}_gpu.txt
DRIVER_HARNESS_KEEP=$PWD/gpurun_out/${TAG}_drivers timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log; tail -2 gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
