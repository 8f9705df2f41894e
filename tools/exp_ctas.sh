for w in cfg4 cfg4s cfg5 cfg3; do for st in 2 3; do for c in 1 2 3 4; do
  echo "$w stages=$st ctas=$c -> $(tools/b.sh --workload $w --steps 60 --stages $st --ctas-per-sm $c)"
done; done; done
