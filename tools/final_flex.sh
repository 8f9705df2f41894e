#!/usr/bin/env bash
# Evidence for the flex kernel (run under gpurun): parity suite, bench lines, launch list, one full ncu capture.
set -u
O=gpurun_out/flexfinal2; mkdir -p $O

for wl in cfg4odd cfg3odd; do
  python bench.py --workload $wl --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench_${wl}_flex.json 2> $O/bench_${wl}.err
done
for wl in cfg4 cfg3 cfg2 cfg5 cfg4s cfg3b cfg3p; do
  python bench.py --workload $wl --family 2 --steps 50 --warmup 5 --no-e2e --no-cpu > $O/bench_${wl}_flex.json 2> $O/bench_${wl}.err
done
python tools/bench_expand.py > $O/expand.txt 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --workload cfg4odd --frames 64"
$CMD > $O/plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg4odd_flex.csv $CMD > $O/ncu_l.log 2>&1
$CMD > $O/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:csic_flex_kernel" -s 4 -c 1 -f -o $O/prof_cfg4odd_flex $CMD > $O/ncu_f.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/flexfinal2/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['roofline']['kernel'], round(d['value']), 'MP/s frac', d['roofline']['frac'], d.get('full_check'), d.get('parity_spot_check'))
    except Exception as e:
        print(f, 'ERR', e)
PY
cat $O/expand.txt
