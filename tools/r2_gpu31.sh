#!/usr/bin/env bash
# pooling planner: CTA sized to small 8x8 tiles; parity + the AVERAGE map
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "average or pooling or randomised or baseline_geometry" > $O/g31_pytest.log 2>&1; tail -3 $O/g31_pytest.log
timeout 900 python tools/perf_map.py CSQ,SQC 1 > $O/g31_perf_map_average.txt 2>&1; tail -3 $O/g31_perf_map_average.txt
timeout 300 python tools/stress.py 30 21 > $O/g31_stress.txt 2>&1; tail -3 $O/g31_stress.txt
