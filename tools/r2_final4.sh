#!/usr/bin/env bash
# definitive closing check of the committed tree: all GPU tests, smoke, the default bench line, decoder, a short stress
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/f4_pytest_gpu.log 2>&1; tail -4 $O/f4_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/f4_smoke.log 2>&1; tail -1 $O/f4_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/f4_bench_cfg4.json 2> $O/f4_bench_cfg4.err; tail -c 200 $O/f4_bench_cfg4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f4_bench_cfg4.json').read().strip().splitlines()[-1]); print(d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_ceiling'), d['value'], d['gpu_launches'])
PY
timeout 300 python tools/stress.py 20 41 > $O/f4_stress.txt 2>&1; tail -3 $O/f4_stress.txt
