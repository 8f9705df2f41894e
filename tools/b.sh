#!/usr/bin/env bash
# one-line summary of a bench run: tools/b.sh <bench args...>
python bench.py --no-e2e --no-cpu "$@" 2>&1 | tail -1 | python -c '
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print("frac", d["roofline"]["frac"], "ms", d["ms_per_step"], "min", d["step_ms_min"], "MP/s", d["value"], "frames", d["config"]["frames_per_gpu"], "parity", d["parity_spot_check"])
except Exception as e:
    print("ERR", e)
'
