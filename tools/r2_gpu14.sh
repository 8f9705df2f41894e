#!/usr/bin/env bash
# TMA decoder: parity first, then timings with tile / stage sweeps
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "decoder or inverse or planar" > $O/g14_pytest_dec.log 2>&1; tail -15 $O/g14_pytest_dec.log
timeout 300 python tools/bench_expand.py > $O/g14_expand_default.txt 2>&1; cat $O/g14_expand_default.txt
for T in 2048 8192; do for S in 2 3 4; do
  echo "== T=$T S=$S"; CSIC_DEC_TILE=$T CSIC_DEC_STAGES=$S timeout 300 python tools/bench_expand.py 2>&1 | tee $O/g14_expand_T${T}_S${S}.txt | cut -c1-90
done; done
for S in 2 4; do echo "== T=4096 S=$S"; CSIC_DEC_TILE=4096 CSIC_DEC_STAGES=$S timeout 300 python tools/bench_expand.py 2>&1 | tee $O/g14_expand_T4096_S${S}.txt | cut -c1-90; done
