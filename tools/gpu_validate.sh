#!/bin/bash
# Round validation on one B200: parity tests, smoke, bench (both arms), ncu launch list, one --set full capture.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e --skip-trial-list > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_score_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --skip-trial-list > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_pairs -s 1 -c 1 -f -o gpurun_out/${TAG}_score_pairs \
    python tools/quick_pairs.py > gpurun_out/${TAG}_ncu_pairs.log 2>&1; echo "ncu pairs rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_grid -s 1 -c 1 -f -o gpurun_out/${TAG}_score_grid \
    python tools/quick_grid.py > gpurun_out/${TAG}_ncu_grid.log 2>&1; echo "ncu grid rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 4 -c 1 -f -o gpurun_out/${TAG}_gemm_tc \
    python tools/quick_train.py > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_train_launches.csv \
    python tools/quick_train.py 131072 > gpurun_out/${TAG}_ncu_train_list.log 2>&1; echo "ncu train list rc=$?"
timeout 300 python tools/quick_train.py > gpurun_out/${TAG}_train.log 2>&1; timeout 300 python tools/quick_grid.py > gpurun_out/${TAG}_grid.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_smoke.log | tail -2; cat gpurun_out/${TAG}_bench.json; cat gpurun_out/${TAG}_bench_ref.json
