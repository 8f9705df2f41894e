#!/usr/bin/env bash
# rows kernel, fused RGB888 with sixteen pixels per trip: residency sweep + one ncu capture
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also --no-verify --workload hd_rgb"
for T in 192 256 384; do for TB in 16384 24576; do for S in 2 3; do
  echo "threads=$T tile=$TB stages=$S $(timeout 200 $B --block-threads $T --tile-bytes $TB --stages $S 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['frac'])")"
done; done; done | tee $O/g20_sweep_hd_rgb.txt
bash tools/ncu_capture.sh rows_hd_rgb16 csic_rows_kernel 3 python bench.py --no-e2e --no-cpu --no-also --workload hd_rgb --frames 64 --steps 2 --warmup 1 --no-verify > $O/g20_ncu.log 2>&1
