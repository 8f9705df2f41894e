#!/usr/bin/env bash
# decoder: third stage where two CTAs per SM still fit; parity (incl. the rows-narrower-than-a-granule cases) and timings against two stages
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "decoder or inverse or planar" > $O/g26_pytest_dec.log 2>&1; tail -5 $O/g26_pytest_dec.log
timeout 300 python tools/bench_expand.py > $O/g26_expand_auto.txt 2>&1; cat $O/g26_expand_auto.txt
CSIC_DEC_STAGES=2 timeout 300 python tools/bench_expand.py > $O/g26_expand_s2.txt 2>&1; cat $O/g26_expand_s2.txt
