#!/usr/bin/env bash
# round 2, third 1-GPU pass: flex kernel v2 + packed RGB clamp: parity, numbers, sweeps, map
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu3.log 2>&1; tail -15 $O/pytest_gpu3.log
for W in wxga_rgb wxga_f2 port_f1 cfg4odd cfg3odd sq200_f4 sq96_f8 hd_rgb hd_f2rgb cfg5; do timeout 300 $B --workload $W > $O/bench_${W}_r2c.json 2>$O/err.txt || tail -3 $O/err.txt; done
for W in cfg4 cfg3 cfg2 cfg5; do timeout 300 $B --workload $W --family 2 > $O/bench_${W}_flex_r2c.json 2>/dev/null; done
timeout 600 python tools/sweep_rows.py hd_rgb 0,2,3,4 2 0,192 0,12288,16384 > $O/sweep_hd_rgb_r2c.txt 2>&1; tail -4 $O/sweep_hd_rgb_r2c.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_flex_kernel -s 3 -c 1 -f -o $O/prof_wxga_rgb_r2c $B --workload wxga_rgb --frames 64 --steps 2 --no-verify > $O/prof_wxga_rgb_r2c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_flex_kernel -s 3 -c 1 -f -o $O/prof_wxga_f2_r2c $B --workload wxga_f2 --frames 64 --steps 2 --no-verify > $O/prof_wxga_f2_r2c.log 2>&1
timeout 600 python tools/perf_map.py > $O/perf_map_r2c.txt 2>&1; tail -12 $O/perf_map_r2c.txt
