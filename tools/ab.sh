# same-box A/B of two library builds: tools/ab.sh <libA.so> <libB.so> workloads...
A=$1; B=$2; shift 2
for w in "$@"; do for rep in 1 2; do for lib in $A $B; do
  r=$(CSIC_LIB_PATH=$lib python bench.py --workload $w --no-e2e --no-cpu --steps 100 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['frac'], d['ms_per_step'], d['step_ms_min'], d['parity_spot_check'])" 2>&1 | tail -1)
  echo "$w $(basename $lib) -> $r"
done; done; done
