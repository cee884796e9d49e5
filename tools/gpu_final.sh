#!/bin/bash
# One GPU-box call at the end of a work block: GPU tests, smoke, both bench workloads (+ reference arm), ncu launch
# list of the small bench, ncu --set full of the HBM-bound kernels (K1 rows, K4), SVM-fit and K1/K4 microbenchmarks.
set -u
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 300 python -m pytest tests/ -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench_ref rc=$?"
timeout 300 python bench.py > $O/bench_eval.json 2> $O/bench_eval.err; echo "bench eval rc=$?"
timeout 300 python bench.py --workload train > $O/bench_train.json 2> $O/bench_train.err; echo "bench train rc=$?"
timeout 60 python tools/bench_k1.py > $O/bench_k1.log 2>&1; echo "bench_k1 rc=$?"
timeout 120 python tools/bench_svm.py > $O/bench_svm.log 2>&1; echo "bench_svm rc=$?"
BENCH_SMALL="python bench.py --steps 1 --warmup 1 --videos-per-step 1 --no-cpu-baseline --no-e2e-jpeg"
timeout 120 $BENCH_SMALL > $O/plain.log 2>&1 &&
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/launches.csv $BENCH_SMALL > $O/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 120 ncu --set full --import-source on --clock-control none -k regex:'preprocess_rows|fuse_kernel' -o $O/k1k4 -f python tools/bench_k1.py --videos 1 --once > $O/ncu_k1k4.log 2>&1
echo "ncu k1k4 rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_after.csv
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; cat $O/bench_k1.log $O/bench_svm.log
