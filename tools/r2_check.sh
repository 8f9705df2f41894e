#!/usr/bin/env bash
# quick 1-GPU re-validation of the tree as committed: GPU tests, smoke, the default bench line, spot workloads
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 900 python -m pytest tests -m gpu -q > $O/check_pytest_gpu.log 2>&1; tail -4 $O/check_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/check_smoke.log 2>&1; tail -1 $O/check_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/check_bench_cfg4.json 2> $O/check_bench_cfg4.err; tail -c 300 $O/check_bench_cfg4.err
for W in cfg3b hd_b128 hd_rgb cfg3 wxga_rgb sq200_f4 cfg4avg; do timeout 300 $B --workload $W > $O/check_bench_${W}.json 2>/dev/null; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/check_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'], (d.get('e2e') or {}).get('value'))
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
