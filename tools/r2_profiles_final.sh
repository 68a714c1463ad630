#!/bin/bash
# Profile refresh at the round's final HEAD (ncu; nothing printed under ncu is a bench value): what changed after
# tools/r2_profiles.sh ran -- the bench step now launches the CTA-pair kernel K1p (score_tcp_kernel) and every form of the
# one-CTA kernel got the barrier probes.  Usage: bash tools/r2_profiles_final.sh <tag>; then here: python tools/summarize_profiles.py <tag> r2
TAG=${1:-r2f}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_launches.out 2>&1; echo "launch list rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_train_launches.csv python tools/ncu_train.py 1000000 d > $O/${TAG}_train.out 2>&1; echo "train launch list rc=$?"
F="--set full --clock-control none --import-source on -f"
timeout 300 ncu $F -k regex:score_tcp_kernel -s 3 -c 1 -o $O/${TAG}_score_tcp python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_k1p.out 2>&1; echo "K1p rc=$?"
# BASELINE configs[2] materialised: 10M pairs = 41 GB through K1p (achieved-GB/s evidence)
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:score_tcp_kernel -s 3 -c 1 --csv --log-file $O/${TAG}_cfg3_10m.csv python bench.py --pairs 10000000 --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_cfg3.out 2>&1; echo "cfg3 rc=$?"
QUICK=1 timeout 300 ncu $F -k regex:score_tcx_kernel -s 2 -c 1 -o $O/${TAG}_score_tcx python tools/quick_split.py 2 > $O/${TAG}_tcx.out 2>&1; echo "K1x rc=$?"
ls -la $O/${TAG}_*
