#!/usr/bin/env bash
# pooling planner, small-CTA rule restricted to tiles under half a CTA: parity + the AVERAGE map once more
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "average or pooling or randomised" > $O/g32_pytest.log 2>&1; tail -3 $O/g32_pytest.log
timeout 900 python tools/perf_map.py CSQ,SQC 1 > $O/g32_perf_map_average.txt 2>&1; tail -3 $O/g32_perf_map_average.txt
