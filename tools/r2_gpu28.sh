#!/usr/bin/env bash
# pooling kernel: held-pair cache keyed by frame and line (bug found by tools/stress.py); regression test, then stress again
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "pooling_first or average" > $O/g28_pytest.log 2>&1; tail -4 $O/g28_pytest.log
timeout 500 python tools/stress.py 120 1 > $O/g28_stress1.txt 2>&1; tail -4 $O/g28_stress1.txt
timeout 500 python tools/stress.py 100 2 > $O/g28_stress2.txt 2>&1; tail -4 $O/g28_stress2.txt
timeout 500 python tools/stress.py 100 3 > $O/g28_stress3.txt 2>&1; tail -4 $O/g28_stress3.txt
