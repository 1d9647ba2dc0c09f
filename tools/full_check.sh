#!/bin/bash
# full GPU validation of a build: the whole -m gpu suite, then bench lines (public bench at the metric's config; learner-only for the others)
TAG=${1:-chk}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_gputest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_5v5.json 2> gpurun_out/${TAG}_bench_5v5.err || tail -5 gpurun_out/${TAG}_bench_5v5.err
python tools/show_bench.py gpurun_out/${TAG}_bench_5v5.json 2>/dev/null | grep -v "^cpu\|hbm:" | cut -c1-170
