#!/bin/bash
# K1p only: launch list of the bench step + one full capture (the rest of tools/r2_profiles_final.sh is unchanged by K1p-only edits).
TAG=${1:-r2L}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_launches.out 2>&1; echo "launch list rc=$?"
F="--set full --clock-control none --import-source on -f"
timeout 300 ncu $F -k regex:score_tcp_kernel -s 3 -c 1 -o $O/${TAG}_score_tcp python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_k1p.out 2>&1; echo "K1p rc=$?"
