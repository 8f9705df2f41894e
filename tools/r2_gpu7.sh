#!/usr/bin/env bash
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also --no-verify"
timeout 600 python -m pytest tests -m gpu -x -q -k "planar or expand or inverse" > $O/pytest_gpu7.log 2>&1; tail -5 $O/pytest_gpu7.log
timeout 300 python tools/bench_expand.py > $O/expand_r2g.txt 2>&1; cat $O/expand_r2g.txt
for W in wxga_rgb port_f1 cfg3odd; do
  $B --workload $W > $O/b_${W}_tall.json 2>/dev/null
  CSIC_FLEX_NO_TALL=1 $B --workload $W > $O/b_${W}_notall.json 2>/dev/null
  for R in 2 3 4 5 6 7 8; do CSIC_FLEX_ROWS=$R $B --workload $W > $O/b_${W}_rows$R.json 2>/dev/null; CSIC_FLEX_NO_TALL=1 CSIC_FLEX_ROWS=$R $B --workload $W > $O/b_${W}_rows${R}_notall.json 2>/dev/null; done
done
for W in wxga_f2 sq200_f4 sq96_f8; do
  for R in 4 7 8 16 32 64; do CSIC_FLEX_ROWS=$R $B --workload $W > $O/b_${W}_rows$R.json 2>/dev/null; done
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
