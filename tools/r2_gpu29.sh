#!/usr/bin/env bash
# stress with the host-path phase (short forward / decoder phases, the host phase is new)
set -x
timeout 400 python tools/stress.py 40 11 > gpurun_out/g29_stress1.txt 2>&1; tail -4 gpurun_out/g29_stress1.txt
timeout 400 python tools/stress.py 40 12 > gpurun_out/g29_stress2.txt 2>&1; tail -4 gpurun_out/g29_stress2.txt
