#!/usr/bin/env bash
set -x
timeout 400 python tools/stress.py 90 1 > gpurun_out/g27_stress1.txt 2>&1; tail -4 gpurun_out/g27_stress1.txt
timeout 400 python tools/stress.py 60 2 > gpurun_out/g27_stress2.txt 2>&1; tail -4 gpurun_out/g27_stress2.txt
