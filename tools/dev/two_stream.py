"""Developer experiment: one batch-64 forward vs two concurrent half-batches on two streams (two plans/workspaces)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nets = [bench.build_model(dev) for _ in range(parts)]
streams = [torch.cuda.Stream() for _ in range(parts)]
sets = [bench.synth_inputs(B, dev, 100 + i) for i in range(4)]
one = bench.build_model(dev)


def run_single(i):
    x, a = sets[i % 4]
    return one(x, a)


def run_split(i):
    x, a = sets[i % 4]
    h = B // parts
    cur = torch.cuda.current_stream()
    outs = []
    for k in range(parts):
        streams[k].wait_stream(cur)
        with torch.cuda.stream(streams[k]):
            outs.append(nets[k](x[k * h:(k + 1) * h], a[k * h:(k + 1) * h]))
    for k in range(parts):
        cur.wait_stream(streams[k])
    return outs


def timeit(fn, n=50):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


a = run_single(0)
b = torch.cat(run_split(0), 0)
torch.cuda.synchronize()
print("split == single:", torch.equal(a, b))
t1 = timeit(run_single)
t2 = timeit(run_split)
print("batch %d: single %.3f ms (%.0f frames/s)   %d streams x %d: %.3f ms (%.0f frames/s)" % (B, t1, B / t1 * 1e3, parts, B // parts, t2, B / t2 * 1e3))
