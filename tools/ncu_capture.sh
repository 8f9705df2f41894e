#!/usr/bin/env bash
# One full ncu capture of one launch, brought back as CSV summaries instead of the (large) .ncu-rep:
#   tools/ncu_capture.sh <name> <kernel-regex> <launch-skip> <command...>
# writes gpurun_out/prof_<name>_raw.csv (all metrics of the launch) and gpurun_out/prof_<name>_source.csv (SASS with
# executed-instruction counts and stall samples); gpurun only copies back 64 MiB.
set -u
NAME=$1; REGEX=$2; SKIP=$3; shift 3
O=gpurun_out
REP=/tmp/prof_${NAME}.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:${REGEX}" -s "$SKIP" -c 1 -f -o "${REP%.ncu-rep}" "$@" > $O/prof_${NAME}.log 2>&1
ncu -i "$REP" --page raw --csv > $O/prof_${NAME}_raw.csv 2>/dev/null
ncu -i "$REP" --page source --csv > $O/prof_${NAME}_source.csv 2>/dev/null
rm -f "$REP"
