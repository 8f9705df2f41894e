#!/usr/bin/env bash
# rows kernel: sixteen pixels per trip for the fused RGB888 reconstruction (f <= 2): parity, benches, RGB column of the perf map
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 900 python -m pytest tests -m gpu -q > $O/g19_pytest_gpu.log 2>&1; tail -4 $O/g19_pytest_gpu.log
for W in hd_rgb hd_f2rgb thumb96rgb cfg4 wxga_rgb; do timeout 300 $B --workload $W > $O/g19_bench_${W}.json 2>/dev/null; done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob('gpurun_out/g19_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(os.path.basename(f), d['roofline']['frac'], d['roofline']['kernel'])
    except Exception as e: print(os.path.basename(f),'FAILED')
PY
timeout 600 python tools/perf_map.py CSQ 0 1 > $O/g19_perf_map_rgb.txt 2>&1; cat $O/g19_perf_map_rgb.txt
