#!/usr/bin/env bash
# row kernel: tall-image mode (tiles span frames) and up to 256 rows per tile; parity + small-frame workloads
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/g23_pytest_gpu.log 2>&1; tail -4 $O/g23_pytest_gpu.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
for V in "notall CSIC_ROWS_NO_TALL=1" "tall64 CSIC_ROWS_MAX_ROWS=64" "tall256 X=1"; do set -- $V
  for W in thumb32 thumb64 thumb128 thumb96rgb thumb256; do
    echo "$1 $W $(env $2 timeout 300 $B --workload $W 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['frac'], d['roofline']['kernel'])")"
  done
done | tee $O/g23_small_frames.txt
for W in cfg4 cfg3 cfg2; do timeout 300 $B --workload $W > $O/g23_bench_${W}.json 2>/dev/null; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/g23_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
