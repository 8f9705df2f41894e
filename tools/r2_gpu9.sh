#!/usr/bin/env bash
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also --no-verify"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu9.log 2>&1; tail -5 $O/pytest_gpu9.log
timeout 300 python tools/bench_expand.py > $O/expand_r2i.txt 2>&1; cat $O/expand_r2i.txt
for W in hd_b128 cfg3b cfg3 cfg3p thumb128 thumb96rgb hd_rgb; do $B --workload $W > $O/bench_${W}_r2i.json 2>/dev/null; done
for T in 0 128 256; do $B --workload thumb96rgb --block-threads $T > $O/bench_thumb96rgb_t$T.json 2>/dev/null; done
bash tools/ncu_capture.sh thumb96rgb_r2i csic_rows_kernel 3 $B --workload thumb96rgb --frames 8192 --steps 2
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/bench_*_r2i.json')+glob.glob('gpurun_out/bench_thumb96rgb_t*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
