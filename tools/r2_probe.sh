#!/bin/bash
# Round-2 hardware probes: gather4 semantics, SS-operand MMA rate, K1 role cycle accounting.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt 2>&1
{
for br in 1 4; do for sw in 0 1 2; do for bc in 64 32; do
  timeout 20 tools/build/gather4_probe $br $sw $bc; echo "rc=$?"
done; done; done
} > gpurun_out/r2a_gather4.log 2>&1
timeout 120 tools/build/ss_probe > gpurun_out/r2a_ss_probe.log 2>&1; echo "ss rc=$?"
timeout 60 tools/build/mma_rate > gpurun_out/r2a_mma_rate.log 2>&1
IMPLS=tc NPLDA_TC_PROF=1 timeout 300 python tools/quick.py 2 > gpurun_out/r2a_k1_prof.log 2>&1; echo "prof rc=$?"
tail -30 gpurun_out/r2a_gather4.log; cat gpurun_out/r2a_ss_probe.log; tail -20 gpurun_out/r2a_k1_prof.log
