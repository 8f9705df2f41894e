# A/B with explicit knobs: tools/ab2.sh <workload> <libA> <libB>
w=$1; A=$2; B=$3
for st in 2 3 4; do for tb in 12288 24576; do for lib in $A $B; do
  r=$(CSIC_LIB_PATH=$lib python bench.py --workload $w --no-e2e --no-cpu --steps 100 --stages $st --tile-bytes $tb 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['frac'], d['ms_per_step'])" 2>&1 | tail -1)
  echo "$w stages=$st tile=$tb $(basename $lib) -> $r"
done; done; done
