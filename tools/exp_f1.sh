for w in cfg3b cfg3p; do for c in 2 3 4; do
  echo "$w ctas=$c -> $(tools/b.sh --workload $w --steps 50 --ctas-per-sm $c)"
done; done
