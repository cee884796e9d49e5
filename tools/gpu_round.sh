#!/bin/bash
# One GPU-box call: GPU test suite, smoke, bench (ours + reference arm), ncu launch list + full capture of the
# tensor-core layer kernel.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
BENCH_SMALL="python bench.py --steps 1 --warmup 1 --videos-per-step 1 --no-cpu-baseline"
$BENCH_SMALL > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv $BENCH_SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BENCH_SMALL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 32 -c 16 -o gpurun_out/prof_conv -f $BENCH_SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_after.csv
