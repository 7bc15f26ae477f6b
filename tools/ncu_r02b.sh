#!/bin/bash
# ncu evidence of the PPO update at the end of round 2 (B200_PROFILING.md recipe): every command first runs plain and must exit 0.
set -u
B="python tools/prof_ppo_update.py"
$B > gpurun_out/r02b_plainB.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_ncu_launches_ppo.csv $B > gpurun_out/r02b_ncuB.log 2>&1
$B > gpurun_out/r02b_plainB2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dw_gemm -s 4 -c 1 -o gpurun_out/r02b_prof_dw_gemm $B > gpurun_out/r02b_ncuB2.log 2>&1
$B > gpurun_out/r02b_plainB3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bptt_persist -s 1 -c 1 -o gpurun_out/r02b_prof_bptt $B > gpurun_out/r02b_ncuB3.log 2>&1
$B > gpurun_out/r02b_plainB4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_save -s 1 -c 1 -o gpurun_out/r02b_prof_fwd_save $B > gpurun_out/r02b_ncuB4.log 2>&1
ls -la gpurun_out/r02b* 2>&1 | tail
