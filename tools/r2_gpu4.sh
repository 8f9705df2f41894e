#!/usr/bin/env bash
# round 2, fourth 1-GPU pass: generic kernel v3, flex sweeps after the restructure, small-frame profiles
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu4.log 2>&1; tail -15 $O/pytest_gpu4.log
for W in cfg4 cfg3 cfg4avg cfg5avg; do timeout 300 $B --workload $W --frames 128 --family 1 --no-verify > $O/bench_${W}_generic_r2d.json 2>/dev/null; done
timeout 300 $B --workload oddavg > $O/bench_oddavg_r2d.json 2>/dev/null
timeout 300 $B --workload thumb32 > $O/bench_thumb32_r2d.json 2>/dev/null
timeout 300 $B --workload thumb64 > $O/bench_thumb64_r2d.json 2>/dev/null
timeout 900 python tools/sweep_flex.py wxga_rgb,wxga_f2,port_f1,cfg4odd,sq200_f4 256 16384,24576,32768,49152 2 > $O/sweep_flex_r2d.txt 2>&1; grep best $O/sweep_flex_r2d.txt
bash tools/ncu_capture.sh sq200_f4_r2d csic_flex_kernel 3 $B --workload sq200_f4 --frames 4096 --steps 2 --no-verify
bash tools/ncu_capture.sh sq96_f8_r2d csic_flex_kernel 3 $B --workload sq96_f8 --frames 16384 --steps 2 --no-verify
bash tools/ncu_capture.sh generic_oddavg_r2d csic_generic_kernel 1 $B --workload oddavg --frames 64 --steps 2 --no-verify
bash tools/ncu_capture.sh generic_cfg4_r2d csic_generic_kernel 1 $B --workload cfg4 --frames 64 --steps 2 --family 1 --no-verify
