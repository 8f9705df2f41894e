#!/usr/bin/env bash
# equal tiles in the row / pooling planners, fill weight in the flex planner: parity + the whole performance map
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/g22_pytest_gpu.log 2>&1; tail -4 $O/g22_pytest_gpu.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
for W in thumb96rgb thumb32 cfg4 cfg4avg oddavg sq200_f4; do timeout 300 $B --workload $W > $O/g22_bench_${W}.json 2>/dev/null; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/g22_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
timeout 900 python tools/perf_map.py CSQ 0 > $O/g22_perf_map.txt 2>&1; tail -12 $O/g22_perf_map.txt
timeout 900 python tools/perf_map.py CSQ,SQC 1 > $O/g22_perf_map_average.txt 2>&1; tail -12 $O/g22_perf_map_average.txt
