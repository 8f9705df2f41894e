#!/usr/bin/env bash
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also --no-verify"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu8.log 2>&1; tail -5 $O/pytest_gpu8.log
for W in wxga_rgb wxga_f2 port_f1 cfg3odd cfg4odd sq200_f4 sq96_f8; do $B --workload $W > $O/bench_${W}_r2h.json 2>/dev/null; done
bash tools/ncu_capture.sh expand_any_r2h csic_expand_planar_any 4 python tools/bench_expand.py
timeout 600 python tools/perf_map.py > $O/perf_map_r2h.txt 2>&1; tail -12 $O/perf_map_r2h.txt
