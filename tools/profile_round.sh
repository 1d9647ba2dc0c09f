#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for one round (run under gpurun; writes gpurun_out/).
#   tools/profile_round.sh [workload] [steps] [full-capture launch count (19 = one step; ~5 s per launch)] [kernel regex of the source-page export] [replay buffer episodes]
#   1. the bench command without ncu (must exit 0)
#   2. launch list with device times (cold-cache, serialised)
#   3. --set full captures of the learner step's kernels; the report stays on the box (gpurun_out/ is capped at 64 MiB):
#      its raw page (all metrics per captured launch) and the source page of the kernels named in $4 come back as CSV
WL=${1:-qmix_5v5_b32}
STEPS=${2:-4}
NFULL=${3:-19}
SRC=${4:-k_gru_}
BUF=${5:-96}
CMD="python bench.py --workload $WL --steps $STEPS --warmup 3 --no-cpu-baseline --learner-only --buffer-size $BUF"
$CMD > gpurun_out/prof_plain_$WL.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain_$WL.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_linear|k_reduce|k_gru|k_agent' -c 900 --csv \
    --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_launches_$WL.log 2>&1
REP=/tmp/prof_round_$WL
ncu --set full --clock-control none --import-source on -k regex:'k_gru_|k_mix_td|k_q_head|k_clip_rmsprop|k_fc2_grad|k_grad_reduce|k_agent_in_tc|k_reduce_tc|k_reduce_group|k_linear_tc|k_linear_group|k_tail' \
    --launch-skip 80 -c $NFULL -o $REP -f $CMD > gpurun_out/ncu_full_$WL.log 2>&1
tail -2 gpurun_out/ncu_full_$WL.log
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/full_raw_$WL.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv -k regex:"$SRC" > gpurun_out/full_source_$WL.csv 2>/dev/null
ls -la gpurun_out/full_*_$WL.csv
