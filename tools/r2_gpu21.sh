#!/usr/bin/env bash
# decoder: per-tile divisions hoisted; parity (incl. the full-size batches) and timings
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "decoder or inverse or planar" > $O/g21_pytest_dec.log 2>&1; tail -5 $O/g21_pytest_dec.log
timeout 300 python tools/bench_expand.py > $O/g21_expand.txt 2>&1; cat $O/g21_expand.txt
