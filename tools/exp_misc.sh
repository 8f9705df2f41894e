for w in cfg3p; do for st in 2 3; do for c in 2 3 4; do
  echo "$w stages=$st ctas=$c -> $(tools/b.sh --workload $w --steps 50 --stages $st --ctas-per-sm $c)"
done; done; done
for tb in 12288 24576 49152; do echo "cfg3p tile=$tb -> $(tools/b.sh --workload cfg3p --steps 50 --tile-bytes $tb)"; done
