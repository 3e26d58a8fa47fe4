#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE GPU, after the plain commands have exited 0):
#   launch list of a whole default bench run, --set full captures of the committed config-5 kernels at full
#   frame, targeted counters of the config-4 kernel at full frame + --set full at 1/16 frame.
O=gpurun_out
M=gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
B="python bench.py --no-extra --no-cpu-baseline --steps 1 --warmup 0"
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_launches_bench.log 2>&1
for prec in f32 hybrid f64; do
  ncu --set full --clock-control none --import-source on -k regex:"k_resolve" -c 1 -f -o $O/r2_c5_${prec} $B --workload c5 --precision $prec > $O/r2_ncu_c5_${prec}.log 2>&1
done
ncu --metrics $M --clock-control none -k regex:k_pt_warp -c 1 --csv --log-file $O/r2_c4_full_frame_counters.csv $B --workload c4 > $O/r2_ncu_c4_counters.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pt_warp -c 1 -f -o $O/r2_c4_scaled4 $B --workload c4 --scale 4 > $O/r2_ncu_c4_s4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pt_warp_bvh -c 1 -f -o $O/r2_c4_bvh_scaled4 $B --workload c4 --scale 4 --accel bvh > $O/r2_ncu_c4_bvh_s4.log 2>&1
ls -la $O
