#!/bin/bash
# Round-2 profile collection on one B200 (ncu; nothing printed under ncu is a bench value).  Usage: bash tools/r2_profiles.sh <tag>
# then here: python tools/summarize_profiles.py <tag> r2
TAG=${1:-r2p}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_launches.out 2>&1; echo "launch list rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_train_launches.csv python tools/ncu_train.py 1000000 d > $O/${TAG}_train.out 2>&1; echo "train launch list rc=$?"
F="--set full --clock-control none --import-source on -f"
timeout 300 ncu $F -k regex:score_tc_kernel -s 3 -c 1 -o $O/${TAG}_score_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_k1.out 2>&1; echo "K1 rc=$?"
# BASELINE configs[2] materialised: 10M pairs = 41 GB through K1 (achieved-GB/s evidence)
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
    --clock-control none -k regex:score_tc_kernel -s 3 -c 1 --csv --log-file $O/${TAG}_cfg3_10m.csv python bench.py --pairs 10000000 --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > $O/${TAG}_cfg3.out 2>&1; echo "cfg3 rc=$?"
# training kernels (262144 pairs: ncu saves and restores everything a kernel writes, GBs per pass at 1M): launch order of tools/ncu_train.py: score_tc_kernel #0 fwd EMIT, #1 guarded fallback, #2 BWD form; then DPlda #3
# (kernel replay ends in LaunchFailed after two passes on these two instantiations -- they write hundreds of MB per launch, which
# ncu saves and restores around every pass -- so they are captured with application replay and the sections that need no patching)
A="--replay-mode application --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --clock-control none -f"
timeout 350 ncu $A -k regex:score_tc_kernel -s 2 -c 1 -o $O/${TAG}_bwd_mid python tools/ncu_train.py 262144 > $O/${TAG}_bwd.out 2>&1; echo "BWD rc=$?"
timeout 350 ncu $A -k regex:score_tc_kernel -s 3 -c 1 -o $O/${TAG}_dplda_fused python tools/ncu_train.py 262144 d > $O/${TAG}_dpl.out 2>&1; echo "DPL rc=$?"
timeout 300 ncu $F -k regex:gemm_tn_tc -s 0 -c 1 -o $O/${TAG}_gemm_tc python tools/ncu_train.py > $O/${TAG}_gemm.out 2>&1; echo "gemm rc=$?"
timeout 300 ncu $F -k regex:wgrad_kernel -c 1 -o $O/${TAG}_dplda_wgrad python tools/ncu_train.py 1000000 d > $O/${TAG}_wgrad.out 2>&1; echo "wgrad rc=$?"
QUICK=1 timeout 300 ncu $F -k regex:score_tcx_kernel -s 2 -c 1 -o $O/${TAG}_score_tcx python tools/quick_split.py 2 > $O/${TAG}_tcx.out 2>&1; echo "K1x rc=$?"
timeout 300 ncu $F -k regex:grid_tc -s 1 -c 1 -o $O/${TAG}_grid_tc python tools/quick_grid.py > $O/${TAG}_grid.out 2>&1; echo "grid rc=$?"
ls -la $O/${TAG}_*
