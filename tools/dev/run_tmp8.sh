cd $GRAFT_REPO_ROOT
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/bench_n$N\_v30.json 2> gpurun_out/bench_n$N\_v30.err
head -c 200 gpurun_out/bench_n$N\_v30.json; echo
done
