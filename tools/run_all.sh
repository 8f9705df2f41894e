#!/usr/bin/env bash
# Round-1 final evidence on one B200 : GPU tests, smoke,
# the driver's own two bench commands, every workload, latency workload with and without a CUDA graph.
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
python bench.py --impl reference 2>&1 | tail -1 > $O/bench_reference_cfg4.json
python bench.py 2>&1 | tail -1 > $O/bench_cfg4.json; cat $O/bench_cfg4.json
for w in cfg4s cfg3 cfg2 cfg5; do python bench.py --workload $w --no-cpu 2>&1 | tail -1 > $O/bench_${w}.json; done
for w in cfg3b cfg3p cfg4avg cfg5avg cfg4savg cfg5savg cfg4odd cfg3odd thumb128 thumb256; do python bench.py --workload $w --no-cpu --no-e2e 2>&1 | tail -1 > $O/bench_${w}.json; done
python bench.py --workload cfg2x1 --steps 200 --no-cpu 2>&1 | tail -1 > $O/bench_cfg2x1_stream.json
python bench.py --workload cfg2x1 --steps 200 --no-cpu --no-e2e --graph 2>&1 | tail -1 > $O/bench_cfg2x1_graph.json
python tools/bench_expand.py > $O/expand.txt 2>&1
python tools/pageable_e2e.py > $O/pageable.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get('e2e') or {}
        print(f.split('/')[-1], d.get('roofline',{}).get('kernel'), round(d['value']), 'MP/s frac', d.get('roofline',{}).get('frac'), 'e2e', e.get('value'), 'ms', d.get('ms_per_step'))
    except Exception as ex:
        print(f, 'ERR', ex)
PY
cat $O/expand.txt $O/pageable.txt
