for w in cfg4avg cfg5avg; do for nt in 256 512; do for c in 2 3 4; do
  echo "$w nt=$nt ctas=$c -> $(tools/b.sh --workload $w --steps 20 --frames ${FR:-256} --block-threads $nt --ctas-per-sm $c)"
done; done; done
