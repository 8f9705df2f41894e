#!/usr/bin/env bash
set -x
timeout 400 python tools/stress.py 60 61 > gpurun_out/g35_stress1.txt 2>&1; tail -4 gpurun_out/g35_stress1.txt
timeout 400 python tools/stress.py 60 62 > gpurun_out/g35_stress2.txt 2>&1; tail -4 gpurun_out/g35_stress2.txt
