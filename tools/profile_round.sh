#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for one round (run under gpurun; writes gpurun_out/).
#   1. the bench command without ncu (must exit 0)
#   2. launch list with device times (cold-cache, serialised)
#   3. --set full captures of the kernels named below
set -e
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_linear|k_reduce|k_gru|k_agent' -c 900 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_gru_fwd4|k_gru_bwd4|k_mix_td|k_q_head|k_record_copy_tma|k_clip_rmsprop|k_fc2_grad|k_grad_reduce|k_agent_in_tc|k_reduce_tc' \
    --launch-skip 40 -c 30 -o gpurun_out/prof_round -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
