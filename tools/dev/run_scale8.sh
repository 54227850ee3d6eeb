cd $GRAFT_REPO_ROOT
nproc; numactl -H 2>/dev/null | head -5; ls /sys/devices/system/node/ | head; cat /sys/devices/system/node/node*/cpulist; python -c "import os; print(len(os.sched_getaffinity(0)))"
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-sweep --no-cpu-baseline > gpurun_out/bench_n$N\_v28.json 2> gpurun_out/bench_n$N\_v28.err
grep "bench\]" gpurun_out/bench_n$N\_v28.err | head -8
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n8_v28.json') if l.startswith('{')][-1])
print(round(d['value']), {k:round(v) for k,v in d['e2e'].items() if k.endswith('value')})
PY
