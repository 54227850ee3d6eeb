#!/usr/bin/env python
"""bench.py -- CASync UNet forward throughput on B200 (BASELINE.json metric: UNet frames/s, 160x160, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one forward pass of the hot path over one batch of synthetic frames per GPU (BASELINE configs[1]:
batch 64, bf16, random-init weights).  Frames are independent, so ranks run disjoint batches with no data-path
collective ("scaling": "weak"); `value` = frames all ranks processed / max-over-ranks device time.

Printed keys (one JSON line, rank 0): the contract's keys plus
  roofline      dominant kernel of the step: algorithmic bytes|flops per launch / CUDA-event time per launch
  cpu_baseline  the oracle port (torch fp32, CPU) timed on this box's host cores on a bounded sample
  e2e           same metric through Model.forward with pinned HOST buffers (H2D + forward + D2H every step)
  stages        per-kernel time split of one profiled step (CUDA events after every launch)
`--impl reference` times the reference algorithm's CPU implementation (the oracle port: the reference itself is
Python and does not exist on the GPU box) with all host threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "unet_frames_per_sec_160x160_bf16", "frames/s"
FLOP_PER_FRAME = 7.9017e9          # 2 x 3.95086 GMAC
STAGEWISE_US_PER_FRAME = 7.11      # sum over stages of max(t_TC, t_HBM) on the measured peaks
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step (BASELINE configs[1] = 64)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons,
                "samples": len(rows)}


def cpu_reference_fps(batch, iters, warmup):
    """The oracle port (reference algorithm, torch fp32 on the host cores).  The one place bench.py runs oracle/."""
    import torch
    from oracle import casync_oracle as O
    sd = O.make_state_dict(0, "R1")
    x, a = O.make_inputs(batch, 0)
    for _ in range(warmup):
        O.forward(sd, x, a)
    ts = []
    for _ in range(iters):
        t = time.perf_counter()
        O.forward(sd, x, a)
        ts.append(time.perf_counter() - t)
    ts.sort()
    return batch / ts[len(ts) // 2], torch.get_num_threads()


def run_reference(args):
    """Reference arm: CPU implementation of the path, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import casync_oracle as O
    try:                                          # torchrun exports OMP_NUM_THREADS=1: use every host thread we may
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    sample = 8                                    # frames per step: a bounded sample of the 64-frame batch
    sd = O.make_state_dict(0, "R1")
    x, a = O.make_inputs(sample, 0)
    for _ in range(args.warmup):
        O.forward(sd, x, a)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.forward(sd, x, a)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    cores = torch.get_num_threads()
    desc = "oracle port of module/unet.py forward, fp32 torch CPU, %d of the %d frames per step" % (sample, args.batch)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "CASync UNet forward, batch %d/GPU, 160x160 (BASELINE configs[1])" % args.batch,
                   "sample_frames_per_step": sample, "host_cpus": os.cpu_count()},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def build_model(device):
    import torch
    from calipsync_b200 import Model
    torch.manual_seed(0)
    net = Model(6, "hubert")                       # random-init weights of the reference architecture
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():                          # healthy BN statistics + live attention path (SURVEY §4 regime R1)
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.75 + 0.5 * torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        for blk in net.attention_blocks:
            blk.cross_attention.gamma.fill_(0.5)
    return net.to(device).eval()


def synth_inputs(batch, device, seed):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.rand(batch, 6, 160, 160, device=device, generator=g)
    x[:, 3:6, 5:150, 5:155] = 0.0                  # the reference's mouth mask (infer_api.py:239)
    a = torch.randn(batch, 32, 32, 32, device=device, generator=g)
    return x, a


def bind_to_gpu_numa_node(local):
    """Pin this rank's threads (and, through first touch, its pinned host buffers) to the NUMA node of its GPU: with 8
    ranks the end-to-end leg moves ~250 GB/s of host memory, which does not fit through the socket interconnect.
    Best effort: says on stderr what it did."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            raise RuntimeError("no NUMA node recorded for %s" % bdf)
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            raise RuntimeError("no allowed CPU on node %d" % node)
        os.sched_setaffinity(0, cpus)
        print("[bench] rank GPU %d (%s): bound to NUMA node %d, %d cpus" % (local, bdf, node, len(cpus)), file=sys.stderr)
    except Exception as e:  # noqa: BLE001
        print("[bench] NUMA binding skipped: %s" % e, file=sys.stderr)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        bind_to_gpu_numa_node(local)
    if world > 1:
        # NCCL announces its version on stdout when the communicator comes up; stdout carries exactly ONE JSON line, so
        # file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    pk, pk_kind = peaks()

    net = build_model(device)
    B = args.batch
    frame_in_bytes = (6 * 160 * 160 + 32 * 32 * 32) * 4
    nsets = max(2, int(1.5 * L2_BYTES / (B * frame_in_bytes)) + 1)   # rotate input sets: total input bytes > L2
    sets = [synth_inputs(B, device, 100 + rank * 1000 + i) for i in range(nsets)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, period=1):
        # untimed prelude before the W warm-up steps: the library captures a CUDA graph the second time it sees the same
        # tensors, so every buffer set is visited twice (eager, capture) before anything counts -- the steady state a
        # caller with a fixed set of buffers is in
        for i in range(2 * period):
            fn(i)
        for i in range(warmup):
            fn(i)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record()
        for i in range(steps):
            fn(i)
        ev1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    # ---- device-resident throughput (value) ------------------------------------------------------------------------
    def step_dev(i):
        x, a = sets[i % nsets]
        net(x, a)

    sampler = ClockSampler(local) if rank == 0 else None
    ms, t0, t1 = timed(step_dev, args.steps, max(3, args.warmup), nsets)
    clocks = sampler.stop(t0, t1) if sampler else None
    fps = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host buffers (e2e) ---------------------------------------------------
    # every step: H2D of that step's fp32 inputs from pinned host memory, Model forward, D2H of its result.
    # (a) sequential on one stream = what infer_api.py:256-266 does; (b) the same three legs of consecutive steps
    # overlapped on three streams by calipsync_b200.HostPipeline (e2e.value); (c) (b) with the uint8 HWC epilogue.
    from calipsync_b200 import HostPipeline
    host_sets = [(x.cpu().pin_memory(), a.cpu().pin_memory()) for x, a in sets[:2]]
    host_out = [torch.empty(B, 3, 160, 160, dtype=torch.float32).pin_memory() for _ in range(2)]
    host_out_u8 = [torch.empty(B, 160, 160, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def step_e2e(i):
        hx, ha = host_sets[i % 2]
        out = net(hx.to(device, non_blocking=True), ha.to(device, non_blocking=True))
        host_out[i % 2].copy_(out, non_blocking=True)

    ms_seq, _, _ = timed(step_e2e, args.steps, max(3, args.warmup), 2)
    fps_e2e_seq = world * B * args.steps / (ms_seq / 1e3)

    def timed_pipeline(pipe, outs, steps, warmup):
        for i in range(warmup + 8):                 # + graph priming: every device slot seen twice (see timed())
            pipe.submit(*host_sets[i % 2], outs[i % 2])
        pipe.flush()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            pipe.submit(*host_sets[i % 2], outs[i % 2])
        ev1.record(pipe.s_out)                      # after the last D2H
        pipe.flush()
        barrier()
        m = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        return float(m.item())

    ms_e2e = timed_pipeline(HostPipeline(net, B), host_out, args.steps, max(3, args.warmup))
    fps_e2e = world * B * args.steps / (ms_e2e / 1e3)
    ms_u8 = timed_pipeline(HostPipeline(net, B, uint8=True), host_out_u8, args.steps, max(3, args.warmup))
    fps_e2e_u8 = world * B * args.steps / (ms_u8 / 1e3)
    # (d) SURVEY 8(f) rows 1+2: the caller's numpy work on the device too -- uint8 crops + frame indices in (HuBERT
    # features of the clip uploaded once), uint8 frames out
    gcpu = torch.Generator().manual_seed(5)
    keep = host_sets
    host_sets = [(torch.randint(0, 256, (B, 160, 160, 3), dtype=torch.uint8, generator=gcpu).pin_memory(),
                  (torch.arange(B, dtype=torch.int32) + 64 * i).pin_memory()) for i in range(2)]
    fpipe = HostPipeline(net, B, frames=True)
    fpipe.set_features(torch.randn(1500, 2, 1024, generator=gcpu).pin_memory())
    ms_fr = timed_pipeline(fpipe, host_out_u8, args.steps, max(3, args.warmup))
    fps_e2e_fr = world * B * args.steps / (ms_fr / 1e3)
    host_sets = keep

    # ---- BASELINE config 3: a 1500-frame clip, strong-scaled (contiguous frame shards), ordered gather INSIDE the timed
    # region: every batch of uint8 frames goes to rank 0 on a side stream while the next batch is computed
    from calipsync_b200 import frame_shard, synthesize_clip
    clip_n = 1500
    clo, chi = frame_shard(clip_n, rank, world)
    gclip = torch.Generator(device=device).manual_seed(4242)            # same features on every rank (replicated)
    clip_feats = torch.randn(clip_n, 2, 1024, device=device, generator=gclip)
    clip_crops = torch.randint(0, 256, (chi - clo, 160, 160, 3), dtype=torch.uint8, device=device,
                               generator=torch.Generator(device=device).manual_seed(5000 + rank))
    clip = {}
    for cb in (64, 256):
        for _ in range(2):                                              # warm-up: communicator, graphs, allocator
            synthesize_clip(net, clip_crops, clip_feats, clip_n, cb)
        barrier()
        reps = 3
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(reps):
            full = synthesize_clip(net, clip_crops, clip_feats, clip_n, cb)
        ev1.record()
        barrier()
        m = torch.tensor([ev0.elapsed_time(ev1)], device=device)
        if world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        clip[str(cb)] = {"ms_per_clip": float(m.item()) / reps, "frames_per_s": clip_n * reps / (float(m.item()) / 1e3)}
        if rank == 0:
            assert full.shape == (clip_n, 160, 160, 3)
    del clip_crops, clip_feats, full

    # ---- BASELINE config 5: batch sweep (device-resident inputs, every rank its own batch, max-over-ranks device time)
    sweep = {}
    if not args.no_sweep:
        for b in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
            try:
                xs = [synth_inputs(b, device, 7 + i) for i in range(2 if b >= 128 else 8)]
                n = max(3, min(40, 8192 // b))
                m, _, _ = timed(lambda i: net(*xs[i % len(xs)]), n, 3, len(xs))
                sweep[str(b)] = round(world * b * n / (m / 1e3), 1)
                del xs
            except Exception as e:      # report, never hide
                sweep[str(b)] = "error: %s" % e
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel split of one step and the roofline of the dominant kernel (rank 0) -----------------------------
    net.profile(*sets[0])
    acc = {}
    reps = 3
    for r in range(reps):
        for rec in net.profile(*sets[(r + 1) % nsets]):
            a = acc.setdefault(rec["name"], dict(ms=0.0, flops=rec["flops"], bytes=rec["bytes"]))
            a["ms"] += rec["ms"] / reps
    total_ms = sum(v["ms"] for v in acc.values())
    hbm_peak, tc_peak = pk["hbm_gbs"], pk["bf16_tflops"]      # burst peaks: each launch is timed between its own events
    stages = []
    for name, v in sorted(acc.items(), key=lambda kv: -kv[1]["ms"]):
        t = v["ms"] / 1e3
        gbs, tfs = v["bytes"] / t / 1e9, v["flops"] / t / 1e12
        bound = "tensor" if v["flops"] / max(v["bytes"], 1) > tc_peak * 1e12 / (hbm_peak * 1e9) else "hbm"
        stages.append({"kernel": name, "ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4), "bound": bound,
                       "gbs": round(gbs, 1), "tflops": round(tfs, 2),
                       "frac": round(tfs / tc_peak if bound == "tensor" else gbs / hbm_peak, 4)})
    # whole step against the per-stage roofline: sum of max(t_TC, t_HBM) over the launches / measured time
    sus = pk.get("bf16_tflops_sustained", tc_peak)
    budget_ms = sum(max(v["flops"] / (sus * 1e12), v["bytes"] / (hbm_peak * 1e9)) for v in acc.values()) * 1e3
    top = stages[0]
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top["kernel"])
    except Exception:
        pass
    roofline = {"kernel": top["kernel"], "bound": top["bound"],
                "achieved": top["tflops"] if top["bound"] == "tensor" else top["gbs"],
                "peak": tc_peak if top["bound"] == "tensor" else hbm_peak,
                "unit": "TFLOP/s" if top["bound"] == "tensor" else "GB/s", "frac": top["frac"], "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json (burst)" if pk_kind == "measured" else "fallback (B200_PROFILING.md)",
                "share_of_step": top["share"],
                "whole_step": {"stagewise_frac": round(STAGEWISE_US_PER_FRAME * 1e-3 * B / (ms / args.steps), 4),
                               "stagewise_budget_ms": round(STAGEWISE_US_PER_FRAME * 1e-3 * B, 4),
                               "launchwise_frac": round(budget_ms / (ms / args.steps), 4),
                               "launchwise_note": "per-launch algorithmic bytes/flops incl. the hidden tensors of the "
                                                  "blocks that are not fused; stagewise = SURVEY 8(d) budget (one "
                                                  "kernel per block, block input + output only)",
                               "tflops": round(fps / world * FLOP_PER_FRAME / 1e12, 1),
                               "frac_of_sustained_bf16": round(fps / world * FLOP_PER_FRAME / 1e12 /
                                                               pk.get("bf16_tflops_sustained", tc_peak), 4)}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, cores = cpu_reference_fps(8, 8, 2)
        v1, _ = cpu_reference_fps(1, 10, 3)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "oracle port (torch fp32 CPU): batch 8 x 8 iters (median); batch 1: %.2f frames/s; host cpus %d"
                         % (v1, os.cpu_count())}

    gpu_ref = None
    if not args.no_cpu_baseline and world == 1:
        # the reference's own deployment is eager PyTorch on the GPU: time the oracle port there too (fp32 and bf16
        # autocast), same batch.  Checker code, not the product path.
        try:
            from oracle import casync_oracle as O
            sd_g = {k: v.to(device) for k, v in net.state_dict().items()}
            xg, ag = sets[0]

            def t_ref(n, autocast):
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    for _ in range(2):
                        O.forward(sd_g, xg, ag)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n):
                        O.forward(sd_g, xg, ag)
                    e1.record()
                    torch.cuda.synchronize()
                return B * n / (e0.elapsed_time(e1) / 1e3)

            gpu_ref = {"what": "oracle port of module/unet.py forward as eager PyTorch on the same B200 (cuDNN/cuBLAS), batch %d" % B,
                       "fp32_frames_per_s": round(t_ref(5, False), 1), "bf16_autocast_frames_per_s": round(t_ref(5, True), 1)}
            gpu_ref["speedup_vs_fp32_eager"] = round(fps / gpu_ref["fp32_frames_per_s"], 2)
            gpu_ref["speedup_vs_bf16_autocast_eager"] = round(fps / gpu_ref["bf16_autocast_frames_per_s"], 2)
            del sd_g
        except Exception as e:          # noqa: BLE001
            gpu_ref = {"error": str(e)}

    line = {
        "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "CASync UNet forward, batch %d/GPU, 160x160 (BASELINE configs[1]), random-init weights" % B,
                   "frames_per_step": world * B, "parallelism": "frame-sharded dp%d, no data-path collective" % world,
                   "l2_policy": "%d rotating input sets (%.0f MB > L2); per-step activations %.1f GB >> L2"
                                % (nsets, nsets * B * frame_in_bytes / 1e6, B * 25e6 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": B * frame_in_bytes, "d2h_bytes_per_step": B * 3 * 160 * 160 * 4,
                "ms_per_step": ms_e2e / args.steps,
                "api": "HostPipeline(Model).submit: fp32 NCHW pinned host in -> Model forward -> fp32 NCHW pinned host out; "
                       "H2D / forward / D2H of consecutive steps overlap on three streams",
                "sequential_value": fps_e2e_seq, "sequential_ms_per_step": ms_seq / args.steps,
                "uint8_out_value": fps_e2e_u8, "uint8_out_d2h_bytes_per_step": B * 160 * 160 * 3,
                "frames_api_value": fps_e2e_fr, "frames_api_h2d_bytes_per_step": B * (160 * 160 * 3 + 4),
                "frames_api": "HostPipeline(frames=True): uint8 crops + frame indices in, uint8 frames out; input "
                              "assembly (infer_api.py:99-145, 238-245) runs on the device",
                "clip_gathered_value": clip["64"]["frames_per_s"], "clip_gathered": clip,
                "clip": "BASELINE configs[2]: 1500-frame clip, contiguous frame shards over %d GPU(s), forward_frames in "
                        "batches of 64 / 256, every batch gathered in order to rank 0 on a side stream (inside the timed "
                        "region, max over ranks); strong scaling" % world},
        "gpu_launches": net.launches_per_forward(B) * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "torch_cuda_reference": gpu_ref, "stages": stages[:12],
        "batch_sweep_frames_per_s": sweep,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
