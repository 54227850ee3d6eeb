cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_n1_v25.json 2> gpurun_out/bench_n1_v25.err; tail -c 600 gpurun_out/bench_n1_v25.json; tail -3 gpurun_out/bench_n1_v25.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01_v25.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/ncu_bench_v25.log 2>&1; tail -2 gpurun_out/ncu_bench_v25.log | cut -c1-300
