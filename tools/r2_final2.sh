#!/usr/bin/env bash
# closing re-validation of the tree on one GPU after the decoder / planner changes: GPU tests, smoke, the default bench
# line (with e2e and the CPU baseline), small-frame and spot workloads, decoder timings, the performance map
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/f2_pytest_gpu.log 2>&1; tail -4 $O/f2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/f2_smoke.log 2>&1; tail -1 $O/f2_smoke.log
timeout 900 python bench.py > $O/f2_bench_cfg4.json 2> $O/f2_bench_cfg4.err; tail -c 300 $O/f2_bench_cfg4.err
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
for W in thumb32 thumb64 thumb128 thumb96rgb cfg3 cfg2 cfg5 cfg3b cfg3p hd_rgb hd_b128 cfg4avg wxga_rgb sq96_f8; do timeout 300 $B --workload $W > $O/f2_bench_${W}.json 2>/dev/null; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/f2_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel'], (d.get('e2e') or {}).get('value'))
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
timeout 300 python tools/bench_expand.py > $O/f2_expand.txt 2>&1; cat $O/f2_expand.txt
timeout 900 python tools/perf_map.py CSQ 0 > $O/f2_perf_map.txt 2>&1; tail -12 $O/f2_perf_map.txt
