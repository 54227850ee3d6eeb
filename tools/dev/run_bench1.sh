cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_n1_v27.json 2> gpurun_out/bench_n1_v27.err; tail -c 300 gpurun_out/bench_n1_v27.json; tail -3 gpurun_out/bench_n1_v27.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v27.json 2> gpurun_out/bench_ref_v27.err; cat gpurun_out/bench_ref_v27.json | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01_v27.csv python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/ncu_bench_v27.log 2>&1; tail -1 gpurun_out/ncu_bench_v27.log | cut -c1-200
