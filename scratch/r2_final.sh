#!/bin/bash
# Round-2 final evidence on ONE GPU: tests, the default bench line, then (after both exited 0) the ncu launch
# list of the same command and the captures of the config-3 kernel.
O=gpurun_out
python -m pytest tests -m gpu -x -q -s > $O/r2_final_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2_final_gputest.log
python bench.py > $O/r2_final_bench_n1.json 2> $O/r2_final_bench_n1.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_final_launches_bench.log 2>&1; echo "launch list rc=$?"
B="python bench.py --no-extra --no-cpu-baseline --steps 1 --warmup 0"
ncu --set full --clock-control none --import-source on -k regex:k_pt_warp -c 1 -f -o $O/r2_final_c3 $B > $O/r2_final_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ncu --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_pt_warp -c 1 --csv --log-file $O/r2_final_c3_flop_counters.csv $B > $O/r2_final_ncu_c3_ops.log 2>&1; echo "ncu ops rc=$?"
