cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_n1_v28.json 2> gpurun_out/bench_n1_v28.err; tail -3 gpurun_out/bench_n1_v28.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1_v28.json'))
print(round(d['value']), round(d['ms_per_step'],4), d['gpu_launches'], {k:round(v) for k,v in d['e2e'].items() if k.endswith('value')}, d['batch_sweep_frames_per_s'], d['clocks'])
PY
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
