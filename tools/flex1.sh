#!/usr/bin/env bash
# flex kernel: parity suite + first numbers
mkdir -p gpurun_out/flex
python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -15 > gpurun_out/flex/pytest.log
cat gpurun_out/flex/pytest.log
for wl in cfg3 cfg4 cfg2 cfg5 cfg4s cfg3b cfg3p; do
  timeout 300 python bench.py --workload $wl --family 2 --steps 30 --warmup 3 --no-e2e --no-cpu --frames 256 > gpurun_out/flex/bench_${wl}_flex.json 2> gpurun_out/flex/bench_${wl}_flex.err
done
for wl in cfg3odd cfg4odd; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --no-e2e --no-cpu > gpurun_out/flex/bench_${wl}.json 2> gpurun_out/flex/bench_${wl}.err
  timeout 300 python bench.py --workload $wl --family 1 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/flex/bench_${wl}_generic.json 2> gpurun_out/flex/bench_${wl}_generic.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/flex/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['roofline']['kernel'], round(d['value']), 'MP/s frac', d['roofline']['frac'], d.get('full_check'), d.get('parity_spot_check'))
    except Exception as e:
        print(f, 'ERR', e)
PY
