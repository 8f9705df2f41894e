#!/usr/bin/env bash
# pooling planner: second row within the slack; experiment: 48 KB tiles for 8x8 pooling
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "average or randomised or baseline_geometry or full_batch" > $O/g24_pytest.log 2>&1; tail -3 $O/g24_pytest.log
timeout 900 python tools/perf_map.py CSQ,SQC 1 > $O/g24_perf_map_average.txt 2>&1; tail -12 $O/g24_perf_map_average.txt
CSIC_POOL_TILE_MAX=49152 timeout 900 python tools/perf_map.py CSQ,SQC 1 > $O/g24_perf_map_average_48k.txt 2>&1; tail -12 $O/g24_perf_map_average_48k.txt
