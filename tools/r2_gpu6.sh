#!/usr/bin/env bash
# round 2, sixth 1-GPU pass: flex planner by residency + tall-image tiles
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu6.log 2>&1; tail -15 $O/pytest_gpu6.log
for W in wxga_rgb wxga_f2 port_f1 cfg4odd cfg3odd sq200_f4 sq96_f8 oddavg; do timeout 300 $B --workload $W > $O/bench_${W}_r2f.json 2>$O/err.txt || tail -3 $O/err.txt; done
for W in cfg4 cfg3 cfg2 cfg5; do timeout 300 $B --workload $W --family 2 > $O/bench_${W}_flex_r2f.json 2>/dev/null; done
timeout 600 python tools/perf_map.py > $O/perf_map_r2f.txt 2>&1; tail -12 $O/perf_map_r2f.txt
