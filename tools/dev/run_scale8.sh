cd $GRAFT_REPO_ROOT
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/bench_n$N\_v25.json 2> gpurun_out/bench_n$N\_v25.err
tail -c 400 gpurun_out/bench_n$N\_v25.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/clip_shard.py 1500 > gpurun_out/clip_shard_n$N.log 2>&1; tail -2 gpurun_out/clip_shard_n$N.log
done
