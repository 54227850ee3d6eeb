#!/bin/bash
# Developer aid: what a round-end check runs on the GPU box (via `gpurun -- bash tools/dev/gpu_check.sh`):
# GPU parity tests, the smoke entry, a short bench line and a per-launch table of one forward at batch 64.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 50 --warmup 5 --no-sweep --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readlines()[-1])
print('bench: %.0f frames/s (%.3f ms/step), e2e %.0f, launches/step %d, roofline %s %.3f' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'] // d['steps'], d['roofline']['kernel'], d['roofline']['frac']))"
[ -x build/casync_run ] && CASYNC_SPLIT=0 CASYNC_GRAPH=0 build/casync_run 64 10 1 2>&1 | tail -75
