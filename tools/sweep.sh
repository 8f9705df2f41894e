for w in cfg4 cfg4s cfg5 cfg3; do
 for tb in 12288 24576 49152; do for st in 2 3 4; do for c in 0 2; do
  r=$(python bench.py --workload $w --no-e2e --no-cpu --steps 30 --frames ${FR:-256} --tile-bytes $tb --stages $st --ctas-per-sm $c 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['frac'], d['ms_per_step'])" 2>&1 | tail -1)
  echo "$w tile=$tb stages=$st ctas=$c -> $r"
 done; done; done
done
