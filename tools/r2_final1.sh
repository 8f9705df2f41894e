#!/usr/bin/env bash
# round 2, closing 1-GPU evidence: tests, smoke, both bench arms, launch list, ncu captures of every kernel, maps
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 900 python -m pytest tests -m gpu -q > $O/final_pytest_gpu.log 2>&1; tail -5 $O/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; tail -2 $O/final_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/final_bench_cfg4.json 2> $O/final_bench_cfg4.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/final_bench_reference.json 2>/dev/null
# launch list of the same command (the step is ONE launch of the row kernel)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches_cfg4.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-also > $O/final_ncu_l.log 2>&1
for W in cfg3 cfg2 cfg5 cfg4s cfg3b cfg3p hd_rgb hd_b128 hd_f2rgb cfg4avg cfg4savg cfg5avg cfg5savg cfg4odd cfg3odd wxga_rgb wxga_f2 port_f1 sq200_f4 sq96_f8 thumb128 thumb64 thumb32 thumb96rgb cfg4f8 cfg4f4 oddavg; do
  timeout 300 $B --workload $W > $O/final_bench_${W}.json 2>/dev/null
done
for W in cfg4 cfg3 cfg4avg; do timeout 300 $B --workload $W --frames 128 --family 1 --no-verify > $O/final_bench_${W}_generic.json 2>/dev/null; done
timeout 300 python tools/bench_expand.py > $O/final_expand.txt 2>&1
NB="$B --no-verify --steps 2"
bash tools/ncu_capture.sh final_rows_cfg4 csic_rows_kernel 3 $NB --workload cfg4 --frames 64
bash tools/ncu_capture.sh final_rows_hd_rgb csic_rows_kernel 3 $NB --workload hd_rgb --frames 64
bash tools/ncu_capture.sh final_rows_hd_b128 csic_rows_kernel 3 $NB --workload hd_b128 --frames 64
bash tools/ncu_capture.sh final_pool_cfg4avg csic_pool_kernel 3 $NB --workload cfg4avg --frames 64
bash tools/ncu_capture.sh final_flex_wxga_rgb csic_flex_kernel 3 $NB --workload wxga_rgb --frames 64
bash tools/ncu_capture.sh final_flex_wxga_f2 csic_flex_kernel 3 $NB --workload wxga_f2 --frames 64
bash tools/ncu_capture.sh final_flex_sq200_f4 csic_flex_kernel 3 $NB --workload sq200_f4 --frames 4096
bash tools/ncu_capture.sh final_generic_cfg4 csic_generic_kernel 1 $NB --workload cfg4 --frames 64 --family 1
bash tools/ncu_capture.sh final_generic_oddavg csic_generic_kernel 1 $NB --workload oddavg --frames 64
bash tools/ncu_capture.sh final_expand_any csic_expand_planar_any 4 python tools/bench_expand.py
timeout 600 python tools/perf_map.py > $O/final_perf_map.txt 2>&1
timeout 600 python tools/perf_map.py SQC > $O/final_perf_map_spatial_first.txt 2>&1
timeout 600 python tools/perf_map.py CSQ,SQC 1 > $O/final_perf_map_average.txt 2>&1
du -sh $O
