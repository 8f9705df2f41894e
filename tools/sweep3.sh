W=${W:-cfg2}
for tb in ${TBS:-12288 24576}; do for st in 2 3 4; do for nt in ${NTS:-256}; do
  r=$(python bench.py --workload $W --no-e2e --no-cpu --steps 100 --tile-bytes $tb --stages $st --block-threads $nt 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['frac'], d['ms_per_step'], d['step_ms_min'])" 2>&1 | tail -1)
  echo "$W nt=$nt tile=$tb stages=$st -> $r"
done; done; done
