#!/bin/bash
# One GPU-box call: ncu evidence for the FINAL layer kernels of one 250-snippet chunk per stream
#   (a) launch list (gpu__time_duration) of the small bench, (b) --set full + tensor-pipe / UTCMMA operand counters of 17
#   layer-kernel launches of the timed step (conv1_1 .. FC3 of the spatial stream, then the temporal stream's conv1_1;
#   layers 2..16 of the two streams are the same kernels on the same shapes).
# Outputs under gpurun_out/evidence/.  The plain run goes first (ncu only after the same command exited 0).
set -u
O=gpurun_out/evidence
mkdir -p $O
BENCH_SMALL="python bench.py --steps 1 --warmup 1 --videos-per-step 1 --no-cpu-baseline --no-e2e-jpeg --no-legs"
timeout 200 $BENCH_SMALL > $O/plain.json 2> $O/plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/launches.csv $BENCH_SMALL > $O/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  MET="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum,l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_2cta.sum,l1tex__data_pipe_tc_wavefronts.sum,sm__cycles_elapsed.max,sm__cycles_active.avg"
  timeout 900 ncu --set full --metrics $MET --clock-control none -k regex:'conv_tc|conv1_fused' -s 32 -c 17 -o $O/prof_layers -f $BENCH_SMALL > $O/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ncu -i $O/prof_layers.ncu-rep --page raw --csv > $O/prof_layers_raw.csv 2> /dev/null
  echo "raw csv lines: $(wc -l < $O/prof_layers_raw.csv)"
  ls -la $O/prof_layers.ncu-rep
  # gpurun_out/ is limited to 64 MiB: the CSV export is what gets committed; keep the report only when it is small
  if [ $(stat -c %s $O/prof_layers.ncu-rep) -gt 30000000 ]; then rm -f $O/prof_layers.ncu-rep; fi
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/smi_after.csv
tail -3 $O/ncu_full.log
