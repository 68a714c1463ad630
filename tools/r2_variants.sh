#!/bin/bash
# A/B of the pair kernel's pipeline variants (experiment builds of csrc/Makefile) + the K1 fp16x3 default.
mkdir -p gpurun_out
L=$PWD/neuralplda_b200
{
echo "== production (NA5 NB4, waits interleaved)"; QUICK=1 timeout 120 python tools/quick_split.py 2 2>&1 | grep "table\|PARITY"
echo "== debug build, variant 0 (interleaved)"; NPLDA_LIB=$L/libnplda_dbg.so NPLDA_TCX_VARIANT=0 QUICK=1 timeout 120 python tools/quick_split.py 2 2>&1 | grep "table"
echo "== debug build, variant 1 (waits after the commits)"; NPLDA_LIB=$L/libnplda_dbg.so NPLDA_TCX_VARIANT=1 QUICK=1 timeout 120 python tools/quick_split.py 2 2>&1 | grep "table"
echo "== NA6 NB3"; NPLDA_LIB=$L/libnplda_a6b3.so QUICK=1 timeout 120 python tools/quick_split.py 2 2>&1 | grep "table"
echo "== NA4 NB4"; NPLDA_LIB=$L/libnplda_a4b4.so QUICK=1 timeout 120 python tools/quick_split.py 2 2>&1 | grep "table"
echo "== K1 (materialised): fp16x3 default vs bf16x3"; IMPLS=tc timeout 200 python tools/quick.py 2 2>&1 | grep -v Warn | tail -14
} > gpurun_out/r2i_variants.log 2>&1
cat gpurun_out/r2i_variants.log
mkdir -p gpurun_out/r2i_drivers
DRIVER_HARNESS_KEEP=$PWD/gpurun_out/r2i_drivers timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; tail -15 gpurun_out/r2i_pytest.log
