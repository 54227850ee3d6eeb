"""Per-launch device-time table of one forward (CUDA events after every launch; see casync_forward_profiled).
    python tools/profile_step.py [batch] [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
net = bench.build_model(dev)
sets = [bench.synth_inputs(B, dev, i) for i in range(3)]
net.profile(*sets[0])
acc, order, reps = {}, [], 5
for r in range(reps):
    for rec in net.profile(*sets[r % 3]):
        if rec["name"] not in acc:
            order.append(rec["name"])
        a = acc.setdefault(rec["name"], dict(ms=0.0, flops=rec["flops"], bytes=rec["bytes"]))
        a["ms"] += rec["ms"] / reps
tot = sum(v["ms"] for v in acc.values())
pk = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
print("batch %d  total %.3f ms  -> %.0f frames/s (event-serialised)" % (B, tot, B / tot * 1e3))
print("%-28s %8s %6s %8s %8s %7s" % ("launch", "ms", "share", "GB/s", "TFLOP/s", "us/frm"))
groups = {}
for n in order:
    v = acc[n]
    t = v["ms"] / 1e3
    print("%-28s %8.4f %5.1f%% %8.0f %8.1f %7.3f" % (n, v["ms"], 100 * v["ms"] / tot, v["bytes"] / t / 1e9,
                                                    v["flops"] / t / 1e12, v["ms"] * 1e3 / B))
    g = n.split(".")[0].rstrip("0123456789") if not n.startswith("attention") else "attention"
    groups[g] = groups.get(g, 0) + v["ms"]
print({k: round(v, 3) for k, v in sorted(groups.items(), key=lambda kv: -kv[1])})
kinds = {}
for n in order:
    k = n.rsplit(".", 1)[-1] if n.rsplit(".", 1)[-1] in ("pw1", "pw2", "dw") else "other"
    kinds[k] = kinds.get(k, 0) + acc[n]["ms"]
print({k: round(v, 3) for k, v in kinds.items()})
if len(sys.argv) > 2:
    json.dump({"batch": B, "total_ms": tot, "launches": [dict(name=n, **acc[n]) for n in order], "peaks": pk},
              open(sys.argv[2], "w"), indent=1)
net.repack()  # destroys the plan (prints developer phase timing when CASYNC_PHASE_DBG is set)
