#!/bin/bash
# Multi-GPU validation (gpurun --gpus N): sharded training check on NCCL, then bench.py at N GPUs with the configs[3] /
# configs[4] legs.  Usage: gpurun --gpus 2 -- 'bash tools/r2_multi.sh 2 r2j'
N=${1:-2}; TAG=${2:-r2multi}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_gpus.txt 2>&1
DIST_BACKEND=nccl DIST_VERBOSE=1 NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29511 tools/dist_train_check.py > gpurun_out/${TAG}_dist_train_nccl.log 2>&1; echo "dist_train rc=$?" | tee -a gpurun_out/${TAG}_dist_train_nccl.log
grep -E "nplda:|dplda:|NCCL INFO (comm|Connected|ncclCommInit).*nranks|rc=" gpurun_out/${TAG}_dist_train_nccl.log | head -20
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py \
    --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench_${N}gpu.json; tail -5 gpurun_out/${TAG}_bench_${N}gpu.err
