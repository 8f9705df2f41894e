#!/usr/bin/env bash
# closing 8-GPU check of the final tree: the two-device test and the bench line at N = 8 (weak scaling + row bands + e2e)
set -x
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "multi or bindings" > $O/final8b_pytest.log 2>&1; tail -3 $O/final8b_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/final8b_bench_n8.json 2> $O/final8b_bench_n8.err
tail -c 400 $O/final8b_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final8b_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('frac_of_ceiling'), d.get('per_rank_ms'))
print({k:(v.get('value'), v.get('speedup_vs_one_gpu')) for k,v in d.get('also',{}).items()})
PY
