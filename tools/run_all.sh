# full round check on one B200: GPU tests, every workload, latency workload with and without a CUDA graph
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
mkdir -p gpurun_out
python bench.py 2>&1 | tail -1 > gpurun_out/bench_cfg4.json; cat gpurun_out/bench_cfg4.json
for w in cfg4s cfg3 cfg2 cfg5; do python bench.py --workload $w 2>&1 | tail -1 | tee gpurun_out/bench_$w.json; done
python bench.py --workload cfg2x1 --steps 200 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_cfg2x1_stream.json
python bench.py --workload cfg2x1 --steps 200 --no-cpu --no-e2e --graph 2>&1 | tail -1 | tee gpurun_out/bench_cfg2x1_graph.json
