# multi-GPU check: tools/run_multi.sh N   (frames sharded weakly for cfg4; row bands of every frame for cfg5)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_cfg4_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 50 --warmup 5 --workload cfg5 --shard bands 2>&1 | tail -1 | tee gpurun_out/bench_cfg5_bands_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 50 --warmup 5 --workload cfg3 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_cfg3_n$N.json
