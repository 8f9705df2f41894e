#!/usr/bin/env bash
# decoder with shared chroma terms: GPU tests, decoder timings, smoke, spot workloads
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 900 python -m pytest tests -m gpu -q > $O/g13_pytest_gpu.log 2>&1; tail -4 $O/g13_pytest_gpu.log
timeout 300 python tools/bench_expand.py > $O/g13_expand.txt 2>&1; cat $O/g13_expand.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/g13_smoke.log 2>&1; tail -1 $O/g13_smoke.log
for W in cfg4 cfg3b hd_rgb; do timeout 300 $B --workload $W > $O/g13_bench_${W}.json 2>/dev/null; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/g13_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
