#!/usr/bin/env bash
# decoder with the unaligned 16-pixel path: parity, timings, small sweep
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "decoder or inverse or planar" > $O/g16_pytest_dec.log 2>&1; tail -5 $O/g16_pytest_dec.log
timeout 300 python tools/bench_expand.py > $O/g16_expand_default.txt 2>&1; cat $O/g16_expand_default.txt
for cfg in "8192 2 256" "4096 3 128" "4096 4 256" "8192 3 128"; do set -- $cfg
  echo "== T=$1 S=$2 NC=$3"; CSIC_DEC_TILE=$1 CSIC_DEC_STAGES=$2 CSIC_DEC_THREADS=$3 timeout 300 python tools/bench_expand.py 2>&1 | tee $O/g16_expand_T$1_S$2_N$3.txt | cut -c1-90
done
