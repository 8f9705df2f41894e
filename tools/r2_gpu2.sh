#!/usr/bin/env bash
# round 2, second 1-GPU pass: parity of the rewritten kernels, then their numbers and ncu summaries
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu2.log 2>&1; tail -15 $O/pytest_gpu2.log
timeout 300 python tools/bench_expand.py > $O/expand_r2.txt 2>&1; cat $O/expand_r2.txt
for W in hd_rgb hd_f2rgb cfg5 cfg4avg cfg4savg cfg5avg cfg5savg oddavg; do timeout 300 $B --workload $W > $O/bench_${W}_r2b.json 2>$O/err.txt || tail -3 $O/err.txt; done
for P in 0 1 2; do timeout 300 $B --workload hd_b128 --store-policy $P > $O/bench_hd_b128_sp${P}.json 2>/dev/null; done
for P in 0 1; do timeout 300 $B --workload cfg3 --store-policy $P > $O/bench_cfg3_sp${P}.json 2>/dev/null; timeout 300 $B --workload cfg4 --store-policy $P > $O/bench_cfg4_sp${P}.json 2>/dev/null; done
# the generic gather kernel forced on BASELINE geometries (what odd case-B shapes and unaligned AVERAGE get)
for W in cfg4 cfg3 cfg4avg; do timeout 300 $B --workload $W --frames 128 --family 1 --no-verify > $O/bench_${W}_generic_r2b.json 2>/dev/null; done
# e2e: chunk size of the host pipeline
for C in 16 32 64 128; do timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-also --no-verify --chunk-mb $C > $O/bench_cfg4_chunk${C}.json 2>/dev/null; done
# ncu summaries of the two rewritten kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_generic_kernel -s 1 -c 1 -f -o $O/prof_generic_oddavg_r2b $B --workload oddavg --frames 64 --steps 2 --no-verify > $O/prof_generic_oddavg.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_generic_kernel -s 1 -c 1 -f -o $O/prof_generic_cfg4_r2b $B --workload cfg4 --frames 64 --steps 2 --family 1 --no-verify > $O/prof_generic_cfg4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_expand_planar_any -s 2 -c 1 -f -o $O/prof_expand_any_r2b python tools/bench_expand.py > $O/prof_expand_any.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_pool_kernel -s 3 -c 1 -f -o $O/prof_cfg4avg_r2b $B --workload cfg4avg --frames 64 --steps 2 --no-verify > $O/prof_cfg4avg_r2b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:csic_rows_kernel -s 3 -c 1 -f -o $O/prof_hd_rgb_r2b $B --workload hd_rgb --frames 64 --steps 2 --no-verify > $O/prof_hd_rgb_r2b.log 2>&1
