#!/bin/bash
# ncu evidence of round 2 (B200_PROFILING.md recipe): every command first runs plain and must exit 0.
set -u
A="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-ppo"
B="python tools/prof_ppo_update.py"
$A > gpurun_out/r02_plainA.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_rollout.csv $A > gpurun_out/r02_ncuA.log 2>&1
$A > gpurun_out/r02_plainA2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_persist -s 2 -c 1 -o gpurun_out/r02_prof_persist $A > gpurun_out/r02_ncuA2.log 2>&1
$B > gpurun_out/r02_plainB.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02_ncu_launches_ppo.csv $B > gpurun_out/r02_ncuB.log 2>&1
$B > gpurun_out/r02_plainB2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bptt_persist -s 1 -c 1 -o gpurun_out/r02_prof_bptt $B > gpurun_out/r02_ncuB2.log 2>&1
$B > gpurun_out/r02_plainB3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_save -s 1 -c 1 -o gpurun_out/r02_prof_fwd_save $B > gpurun_out/r02_ncuB3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_ncu_launches_*.csv 2>&1 | tail
