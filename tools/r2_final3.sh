#!/usr/bin/env bash
# last validation of the tree: all GPU tests, smoke, decoder timings with the final stage rule, the spatial-first map, default bench line
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/f3_pytest_gpu.log 2>&1; tail -4 $O/f3_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/f3_smoke.log 2>&1; tail -1 $O/f3_smoke.log
timeout 300 python tools/bench_expand.py > $O/f3_expand.txt 2>&1; cat $O/f3_expand.txt
timeout 900 python tools/perf_map.py SQC 0 > $O/f3_perf_map_spatial_first.txt 2>&1; tail -12 $O/f3_perf_map_spatial_first.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/f3_bench_cfg4.json 2> $O/f3_bench_cfg4.err; tail -c 200 $O/f3_bench_cfg4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f3_bench_cfg4.json').read().strip().splitlines()[-1]); print(d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_ceiling'), d['value'])
PY
