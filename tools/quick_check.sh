#!/bin/bash
# quick GPU check of a kernel change: parity tests of the learner / GEMMs, then learner-only bench lines (per-kernel table)
#   tools/quick_check.sh [tag] [workloads...]
TAG=${1:-chk}; shift
WLS=${@:-"qmix_20v20_b1024 qmix_10v10_b128 qmix_5v5_b32"}
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_learner.py -x -q -m gpu 2>&1 | tail -5
for WL in $WLS; do
  timeout 300 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --learner-only --buffer-size 96 > gpurun_out/${TAG}_learner_$WL.json 2> gpurun_out/${TAG}_learner_$WL.err || tail -5 gpurun_out/${TAG}_learner_$WL.err
  python tools/show_bench.py gpurun_out/${TAG}_learner_$WL.json 2>/dev/null | grep -v "^cpu\|hbm:" | cut -c1-110
done
