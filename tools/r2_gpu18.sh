#!/usr/bin/env bash
# full validation with the TMA decoder + ncu captures of the decoder
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/g18_pytest_gpu.log 2>&1; tail -4 $O/g18_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/g18_smoke.log 2>&1; tail -1 $O/g18_smoke.log
timeout 300 python tools/bench_expand.py > $O/g18_expand.txt 2>&1; cat $O/g18_expand.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/g18_bench_cfg4.json 2> $O/g18_bench_cfg4.err; tail -c 300 $O/g18_bench_cfg4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g18_bench_cfg4.json').read().strip().splitlines()[-1]); print(d['roofline']['frac'], d['e2e']['value'], d['value'])
PY
# ncu: the decoder's dominant launches (1080p 4:2:0 RGB = 2nd expand launch group; 1918x1078 RGB)
bash tools/ncu_capture.sh decode_hd_rgb csic_decode_kernel 9 python tools/bench_expand.py > $O/g18_ncu1.log 2>&1; tail -2 $O/g18_ncu1.log
bash tools/ncu_capture.sh decode_odd_rgb csic_decode_kernel 37 python tools/bench_expand.py > $O/g18_ncu2.log 2>&1; tail -2 $O/g18_ncu2.log
