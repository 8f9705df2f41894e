#!/usr/bin/env bash
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also --no-verify"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu12.log 2>&1; tail -5 $O/pytest_gpu12.log
for W in sq200_f4 sq96_f8 wxga_rgb wxga_f2 port_f1 cfg3odd cfg4odd; do $B --workload $W > $O/bench_${W}_r2l.json 2>/dev/null; done
timeout 600 python tools/perf_map.py > $O/perf_map_r2l.txt 2>&1
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/bench_*_r2l.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
