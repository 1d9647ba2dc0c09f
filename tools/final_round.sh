#!/bin/bash
# end-of-session evidence run (under gpurun): GPU suite, public bench at the metric's config + reference arm, the other configs
TAG=${1:-r02f}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${TAG}_gputest.log; tail -2 gpurun_out/${TAG}_gputest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_5v5.json 2> gpurun_out/${TAG}_bench_5v5.err || tail -5 gpurun_out/${TAG}_bench_5v5.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null
for WL in qmix_10v10_b128 qmix_20v20_b1024 vdn_5v5_b32; do
  timeout 900 python bench.py --workload $WL --no-cpu-baseline > gpurun_out/${TAG}_bench_$WL.json 2> gpurun_out/${TAG}_bench_$WL.err || tail -5 gpurun_out/${TAG}_bench_$WL.err
done
for f in gpurun_out/${TAG}_bench_5v5.json gpurun_out/${TAG}_bench_qmix_10v10_b128.json gpurun_out/${TAG}_bench_qmix_20v20_b1024.json gpurun_out/${TAG}_bench_vdn_5v5_b32.json; do
  python tools/show_bench.py $f 2>/dev/null | grep -v "^cpu\|hbm:\|act_select" | head -8 | cut -c1-150
done
