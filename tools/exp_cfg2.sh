for a in "--stages 2 --ctas-per-sm 2" "--stages 2 --ctas-per-sm 3" "--stages 3 --ctas-per-sm 2" "--stages 2 --ctas-per-sm 2 --block-threads 128" "--stages 3 --ctas-per-sm 2 --block-threads 128" "--stages 3 --ctas-per-sm 3 --block-threads 128 --tile-bytes 12288"; do
  echo "cfg2 $a -> $(tools/b.sh --workload cfg2 --steps 100 $a)"
done
for fr in 1 4 16; do echo "cfg2 frames=$fr -> $(tools/b.sh --workload cfg2 --frames $fr --steps 200)"; done
