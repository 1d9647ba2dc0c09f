set -x
CMD="python bench.py --workload qmix_20v20_b1024 --steps 2 --warmup 3 --no-cpu-baseline --learner-only --buffer-size 96"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_agent_in_tc|k_linear_tc2' --launch-skip 20 -c 5 -o /tmp/ai -f $CMD > gpurun_out/ncu_ai.log 2>&1
tail -3 gpurun_out/ncu_ai.log
ncu -i /tmp/ai.ncu-rep --page raw --csv > gpurun_out/ai_raw.csv 2>/dev/null
ncu -i /tmp/ai.ncu-rep --page source --csv -k regex:k_agent_in_tc > gpurun_out/ai_source.csv 2>/dev/null
ncu -i /tmp/ai.ncu-rep --page source --csv -k regex:k_linear_tc2 > gpurun_out/lin2_source.csv 2>/dev/null
ls -la gpurun_out/ai_*.csv gpurun_out/lin2_source.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_agent_step|k_rollout|k_dqn|k_eps' -c 400 --csv --log-file gpurun_out/launches_actsel.csv python tools/actsel_bench.py > gpurun_out/ncu_actsel.log 2>&1
tail -5 gpurun_out/ncu_actsel.log
