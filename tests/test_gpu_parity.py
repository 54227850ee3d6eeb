"""Parity of the CUDA path (through the drop-in Model -> C ABI) against the oracle and the golden vectors.

Tolerances are BASELINE.json's: bf16 path max-abs <= 2/255 per pixel and PSNR >= 45 dB (peak 1.0) against the
reference fp32 forward; per-stage relative L2 <= 2e-2 (bf16 operands: ~3e-3 expected, packed_emulation.py).
"""
import os

import numpy as np
import pytest
import torch

from calipsync_b200 import Model
from oracle import casync_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAX_ABS_255, MIN_PSNR, STAGE_REL = 2.0, 45.0, 2e-2
_HW = {"x1": 160, "x2": 80, "x3": 40, "x4": 20, "up1": 20, "up2": 40, "up3": 80, "up4": 160}


def make_model(regime, seed=0):
    sd = O.make_state_dict(seed, regime)
    m = Model(6, "hubert")
    m.load_state_dict(sd, strict=True)
    return m.to("cuda:0").eval(), sd


def stage_nchw(model, name, batch):
    v = model.stage(name, batch).float().cpu()
    hw = _HW.get(name, 10)
    return v.reshape(batch, hw, hw, v.shape[1]).permute(0, 3, 1, 2)


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_forward_matches_golden_and_oracle_stages(regime, monkeypatch):
    monkeypatch.setenv("CASYNC_FUSE_OUTC", "0")      # keep stage "up4" materialised (fused head: test below)
    model, sd = make_model(regime)
    x, a = O.make_inputs(2, 0)
    out = model(x.cuda(), a.cuda()).cpu()
    assert out.shape == (2, 3, 160, 160) and out.dtype == torch.float32
    ref, rst = O.forward(sd, x, a, return_stages=True)
    report = []
    for name in O.STAGE_NAMES[:-2]:
        report.append((name, O.rel_l2(stage_nchw(model, name, 2), rst[name])))
    print("\n[%s] per-stage rel-L2: " % regime + "  ".join("%s=%.4f" % r for r in report))
    print("[%s] out: max-abs*255=%.4f PSNR=%.2f dB" % (regime, O.max_abs_255(out, ref), O.psnr_db(out, ref)))
    bad = [r for r in report if not r[1] < STAGE_REL]
    assert not bad, bad
    gold = torch.from_numpy(np.load(os.path.join(GOLD, "unet_%s_seed0_b2.npz" % regime))["out"])
    for tgt in (ref, gold):          # oracle here, and the reference's own recorded output
        assert O.max_abs_255(out, tgt) <= MAX_ABS_255
        assert O.psnr_db(out, tgt) >= MIN_PSNR
        for i in range(2):
            assert O.psnr_db(out[i], tgt[i]) >= MIN_PSNR


@pytest.mark.parametrize("inctc", ["1", "0"])
@pytest.mark.parametrize("batch", [3, 37])
def test_input_block_on_tensor_cores(batch, inctc, monkeypatch):
    """InConvDw (module/unet.py:58-67) as a strip_tc instantiation (fp32 input as bf16 hi + lo, biases on constant-one
    channels, all three convolutions on tcgen05) and as round 1's inc_kernel (CASYNC_INCTC=0): stage x1 against the oracle.
    Batches that are no multiple of anything: the CTAs' row ranges start and end inside frames and strips."""
    monkeypatch.setenv("CASYNC_INCTC", inctc)
    monkeypatch.setenv("CASYNC_SPLIT", "0")
    model, sd = make_model("R1", seed=3)
    x, a = O.make_inputs(batch, 5)
    model(x.cuda(), a.cuda())
    ref = torch.cat([O.forward(sd, x[i:i + 8], a[i:i + 8], return_stages=True)[1]["x1"] for i in range(0, batch, 8)])
    err = O.rel_l2(stage_nchw(model, "x1", batch), ref)
    print("\n[inc, INCTC=%s, B=%d] x1 rel-L2 = %.5f" % (inctc, batch, err))
    assert err < 6e-3, err


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_config2_batch64_tolerance(regime):
    """BASELINE config 2: batch 64, bf16 on one B200 vs the fp32 reference arithmetic (oracle), default-init (R0) and
    perturbed (R1) weights."""
    model, sd = make_model(regime, seed=1)
    x, a = O.make_inputs(64, 11)
    out = model(x.cuda(), a.cuda()).cpu()
    ref = torch.cat([O.forward(sd, x[i:i + 16], a[i:i + 16]) for i in range(0, 64, 16)])
    print("\n[B=64] max-abs*255=%.4f PSNR=%.2f dB" % (O.max_abs_255(out, ref), O.psnr_db(out, ref)))
    assert O.max_abs_255(out, ref) <= MAX_ABS_255 and O.psnr_db(out, ref) >= MIN_PSNR
    assert min(O.psnr_db(out[i], ref[i]) for i in range(64)) >= MIN_PSNR


@pytest.mark.parametrize("batch", [3, 40])
def test_output_head_fused_into_the_last_decoder_block_is_bit_exact(batch, monkeypatch):
    """OutConv + outc_bn + sigmoid (module/unet.py:342-344) run in the epilogue of up4.1 (no outc launch, the block's
    bf16 output never reaches HBM).  The head is evaluated on the ROUNDED bf16 activations in the order of the stand-alone
    kernel, so fp32 and uint8 outputs must equal the two-launch path bit for bit."""
    x, a = O.make_inputs(batch, 15)
    xs, as_ = x.cuda(), a.cuda()
    fused, sd = make_model("R1", seed=8)
    out_f, u8_f = fused(xs, as_), fused.forward_uint8(xs, as_)
    monkeypatch.setenv("CASYNC_FUSE_OUTC", "0")
    plain, _ = make_model("R1", seed=8)
    assert torch.equal(out_f, plain(xs, as_))
    assert torch.equal(u8_f, plain.forward_uint8(xs, as_))
    assert fused.launches_per_forward(batch) == plain.launches_per_forward(batch) - 1   # (split batches: the tail runs once)
    ref = O.forward(sd, x, a)
    assert O.max_abs_255(out_f.cpu(), ref) <= MAX_ABS_255 and O.psnr_db(out_f.cpu(), ref) >= MIN_PSNR


@pytest.mark.parametrize("batch", [5, 33])
def test_no_writes_outside_the_callers_buffers(batch):
    """Canary regions around the output tensor and the workspace handed to casync_forward (C ABI, unsplit and two-lane
    batches, fp32 and uint8 outputs) must stay untouched, and results must not depend on what the workspace held."""
    import ctypes
    from calipsync_b200 import _lib
    model, _ = make_model("R1", seed=9)
    x, a = O.make_inputs(batch, 17)
    xs, as_ = x.cuda(), a.cuda()
    want = model(xs, as_)                                  # creates the plan
    lib, plan = _lib.load(), model._plan[0]
    pad = 1 << 20
    ws_bytes = lib.casync_workspace_bytes(plan, batch)
    ws = torch.full((ws_bytes + 2 * pad,), 0xA5, dtype=torch.uint8, device="cuda")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for flags, frame_bytes in ((_lib.F_BF16, 3 * 160 * 160 * 4), (_lib.F_OUT_U8_HWC, 160 * 160 * 3)):
        out = torch.full((batch * frame_bytes + 2 * pad,), 0x5A, dtype=torch.uint8, device="cuda")
        assert (ws.data_ptr() + pad) % 256 == 0 and (out.data_ptr() + pad) % 256 == 0
        rc = lib.casync_forward(plan, xs.data_ptr(), as_.data_ptr(), out.data_ptr() + pad, ws.data_ptr() + pad, batch, flags,
                                stream)
        _lib.check(rc, "casync_forward")
        torch.cuda.synchronize()
        assert bool((out[:pad] == 0x5A).all()) and bool((out[-pad:] == 0x5A).all())
        assert bool((ws[:pad] == 0xA5).all()) and bool((ws[-pad:] == 0xA5).all())
        body = out[pad: pad + batch * frame_bytes]
        if flags == _lib.F_BF16:
            assert torch.equal(body.view(torch.float32).view(batch, 3, 160, 160), want)
        else:
            assert torch.equal(body.view(batch, 160, 160, 3), model.forward_uint8(xs, as_))


def test_decoder_first_block_paths_agree(monkeypatch):
    """up1.0 / up2.0 run as upsample-concat pass + TMA-fed GEMM + depthwise + GEMM (default).  The older paths -- concat
    gathered in the GEMM's A producer (fp32 taps), up2.0 on the weight-streaming fused kernel -- stay selectable and must
    give the same image within bf16 noise, each within the tolerance against the oracle."""
    x, a = O.make_inputs(6, 23)
    new, sd = make_model("R1", seed=10)
    out_new = new(x.cuda(), a.cuda()).cpu()
    monkeypatch.setenv("CASYNC_UPCAT_PASS", "0")
    monkeypatch.setenv("CASYNC_UNFUSE_UP2", "0")
    old, _ = make_model("R1", seed=10)
    out_old = old(x.cuda(), a.cuda()).cpu()
    assert new.launches_per_forward(6) == old.launches_per_forward(6) + 4      # + 2 passes, up2.0: 1 -> 3 launches
    ref = O.forward(sd, x, a)
    for out in (out_new, out_old):
        assert O.max_abs_255(out, ref) <= MAX_ABS_255 and O.psnr_db(out, ref) >= MIN_PSNR
    assert O.max_abs_255(out_new, out_old) <= 0.5 and O.psnr_db(out_new, out_old) >= 60.0


def test_frames_are_independent_and_ragged_batches_work():
    """Any batch size (not a multiple of the 128-row tiles), and frame i does not depend on its batch mates:
    the property frame sharding relies on (bit-exact, same kernels and per-row arithmetic)."""
    model, _ = make_model("R1")
    x, a = O.make_inputs(7, 3)
    xg, ag = x.cuda(), a.cuda()
    full = model(xg, ag)
    for lo, hi in ((0, 1), (1, 4), (4, 7), (6, 7)):
        part = model(xg[lo:hi], ag[lo:hi])
        assert torch.equal(part, full[lo:hi]), (lo, hi)
    assert torch.equal(model(xg, ag), full)      # deterministic


def test_chunked_execution_equals_single_pass(monkeypatch):
    model, _ = make_model("R1")
    x, a = O.make_inputs(10, 4)
    full = model(x.cuda(), a.cuda())
    monkeypatch.setenv("CASYNC_CHUNK", "4")       # 10 frames -> passes of 4, 4, 2
    model.repack()
    assert torch.equal(model(x.cuda(), a.cuda()), full)
    monkeypatch.delenv("CASYNC_CHUNK")
    model.repack()


def test_uint8_hwc_epilogue_truncates_like_the_caller():
    """infer_api.py:265-266: pred.cpu().numpy().transpose(1,2,0)*255 -> np.uint8 (truncation)."""
    model, _ = make_model("R1")
    x, a = O.make_inputs(3, 9)
    f = model(x.cuda(), a.cuda()).cpu()
    u = model.forward_uint8(x.cuda(), a.cuda()).cpu()
    assert u.shape == (3, 160, 160, 3) and u.dtype == torch.uint8
    expect = np.array(f.numpy().transpose(0, 2, 3, 1) * 255, dtype=np.uint8)
    # byte work: the kernel truncates the very fp32 product the caller computes -> exact, not "close"
    assert np.array_equal(u.numpy(), expect)


def test_predictions_for_the_unmodified_callers_batch():
    """tests/golden/caller_batch.npz was recorded from the reference's own FrameSynthesizer.process_batch (reference
    Model, CPU): the crops it cut, the HuBERT windows it built and the uint8 frames it pasted back.  Fed with exactly
    those inputs the CUDA path must give the caller the same bytes within 2/255 -- through forward() + the caller's own
    `* 255 -> uint8` and through the uint8 epilogue."""
    gold = np.load(os.path.join(GOLD, "caller_batch.npz"))
    model, _ = make_model("R1", seed=5)
    x = O.assemble_x(gold["crops"]).cuda()
    a = torch.from_numpy(gold["audio"]).cuda()
    pred = model(x, a)
    got = np.stack([np.array(pred[i].cpu().numpy().transpose(1, 2, 0) * 255, dtype=np.uint8) for i in range(len(pred))])
    diff = np.abs(got.astype(np.int32) - gold["pred_u8"].astype(np.int32))
    print("\n[caller batch] max |byte diff| = %d, differing bytes %.2f %%" % (diff.max(), 100 * (diff != 0).mean()))
    assert diff.max() <= 2
    assert np.array_equal(model.forward_uint8(x, a).cpu().numpy(), got)
    u = model.forward_frames(torch.from_numpy(gold["crops"]).cuda(), torch.zeros(1, 2, 1024).cuda(),
                             torch.zeros(len(pred), dtype=torch.int32).cuda())     # shape / dtype contract of the frames API
    assert u.shape == got.shape and u.dtype == torch.uint8


@pytest.mark.parametrize("soft", [False, True])
def test_device_blend_is_bit_exact_with_numpy(soft):
    """casync_blend_paste against the oracle's restatement of the caller's float64 blend: byte
    work, so every frame must match exactly -- ragged regions of different sizes per frame, hard 0/255 polygon masks,
    and the optional soft per-frame mask."""
    from calipsync_b200 import blend_paste
    rs = np.random.RandomState(17)
    b, h, w, ldc = 5, 300, 280, 200
    frames = rs.randint(0, 256, size=(b, h, w, 3), dtype=np.uint8)
    crops = rs.randint(0, 256, size=(b, ldc, ldc, 3), dtype=np.uint8)
    face = (rs.rand(b, ldc, ldc) > 0.4).astype(np.uint8) * 255
    face[0, :, :7] = 128                                                  # not only 0 / 255: m = 128/255
    softm = rs.rand(b, ldc, ldc).astype(np.float32) if soft else None
    rects = np.array([[10, 10 + 200, 5, 5 + 200], [0, 150, 0, 150], [100, 300, 80, 280], [37, 38, 11, 12],
                      [50, 50 + 173, 60, 60 + 173]], dtype=np.int32)
    want = np.stack([O.blend_paste(frames[i], crops[i, : rects[i, 1] - rects[i, 0], : rects[i, 3] - rects[i, 2]],
                                   face[i, : rects[i, 1] - rects[i, 0], : rects[i, 3] - rects[i, 2]], tuple(rects[i]),
                                   None if softm is None else softm[i, : rects[i, 1] - rects[i, 0], : rects[i, 3] - rects[i, 2]])
                     for i in range(b)])
    got = blend_paste(torch.from_numpy(frames.copy()).cuda(), torch.from_numpy(crops).cuda(), torch.from_numpy(face).cuda(),
                      torch.from_numpy(rects).cuda(), None if softm is None else torch.from_numpy(softm).cuda())
    assert np.array_equal(got.cpu().numpy(), want)


def test_three_argument_convenience_and_input_preservation():
    model, _ = make_model("R0")
    x, a = O.make_inputs(2, 5)
    xg, ag = x.cuda(), a.cuda()
    out = model(xg, ag)
    assert torch.equal(xg.cpu(), x) and torch.equal(ag.cpu(), a)          # inputs not mutated
    assert torch.equal(model(xg[:, :3], xg[:, 3:], ag), out)
    assert torch.equal(model(xg.flip(0).flip(0), ag), out)


def test_errors_surface_as_runtime_errors():
    model, _ = make_model("R0")
    x, a = O.make_inputs(1, 0)
    with pytest.raises(RuntimeError):
        model(x.cuda(), a)                       # mixed devices
    with pytest.raises(RuntimeError):
        model(x.cuda()[:, :, :80], a.cuda())     # wrong spatial size
    model.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        model(x.cuda(), a.cuda())


def test_called_from_worker_thread_on_side_stream():
    """Streaming mode calls the model from a non-main thread (image_infer_v1/infer_api.py:193-194)."""
    import threading
    model, _ = make_model("R1")
    x, a = O.make_inputs(2, 6)
    want = model(x.cuda(), a.cuda())
    got = {}

    def work():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            got["out"] = model(x.cuda(), a.cuda())
        s.synchronize()

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert torch.equal(got["out"], want)


def test_two_threads_sharing_one_model_serialise():
    """Two Python threads calling ONE Model on their own streams (ctypes releases the GIL): calls serialise on the
    model's lock and on a completion event, so both get the single-threaded result (the reference nn.Module is safe
    under this use; plan, lanes and workspace are shared state here)."""
    import threading
    model, _ = make_model("R1", seed=6)
    ins = [O.make_inputs(24 + i, 40 + i) for i in range(2)]         # >= 24 frames: two-lane split, plan-owned streams
    want = [model(x.cuda(), a.cuda()).clone() for x, a in ins]
    got = [[None] * 6, [None] * 6]
    errs = []

    def work(i):
        try:
            s = torch.cuda.Stream()
            xg, ag = ins[i][0].cuda(), ins[i][1].cuda()
            torch.cuda.synchronize()
            with torch.cuda.stream(s):
                for r in range(6):
                    got[i][r] = model(xg, ag)
            s.synchronize()
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(2):
        for r in range(6):
            assert torch.equal(got[i][r], want[i]), (i, r)


def test_unfused_switch_runs_every_block_through_the_gemm_path(monkeypatch):
    """CASYNC_NO_FUSED_IR=1 (INTEGRATION.md A/B switch): every InvertedResidual as pw1 GEMM + depthwise + pw2 GEMM,
    including the 32-channel blocks whose K is half a 64-channel k-block.  Same tolerance as the default path."""
    monkeypatch.setenv("CASYNC_NO_FUSED_IR", "1")
    model, sd = make_model("R1", seed=7)
    x, a = O.make_inputs(3, 13)
    out = model(x.cuda(), a.cuda()).cpu()
    ref = O.forward(sd, x, a)
    assert O.max_abs_255(out, ref) <= MAX_ABS_255 and O.psnr_db(out, ref) >= MIN_PSNR


def test_host_pipeline_equals_direct_forward():
    """HostPipeline (H2D / forward / D2H of consecutive batches on three streams) returns exactly what a direct
    Model call returns, for more batches than slots, a ragged last batch, and the uint8 epilogue."""
    from calipsync_b200 import HostPipeline
    model, _ = make_model("R1")
    batches = [O.make_inputs(n, 20 + i) for i, n in enumerate((4, 4, 4, 4, 3))]
    want = [model(x.cuda(), a.cuda()).cpu() for x, a in batches]
    want_u8 = [model.forward_uint8(x.cuda(), a.cuda()).cpu() for x, a in batches]
    for uint8, ref in ((False, want), (True, want_u8)):
        pipe = HostPipeline(model, 4, uint8=uint8)
        outs = [torch.empty_like(r).pin_memory() for r in ref]
        for (x, a), o in zip(batches, outs):
            pipe.submit(x.pin_memory(), a.pin_memory(), o)
        pipe.flush()
        for o, r in zip(outs, ref):
            assert torch.equal(o, r)


def test_device_input_prologue_is_bit_exact_and_feeds_the_forward():
    """casync_prepare_inputs (uint8 crops -> 6-channel tensor, HuBERT window gather) against the oracle's restatement
    of the caller's numpy code -- integer/byte work and exact float32 divisions: bit-exact, incl. clip-end zero rows."""
    model, _ = make_model("R1")
    rs = np.random.RandomState(5)
    crops = torch.from_numpy(rs.randint(0, 256, size=(5, 160, 160, 3), dtype=np.uint8))
    feats = torch.from_numpy(rs.randn(37, 2, 1024).astype(np.float32))
    idxs = [0, 3, 18, 36, 38]                         # clip start, interior, clip end, past the end (reference: zeros)
    x, a = model.prepare_inputs(crops.cuda(), feats.cuda(), torch.tensor(idxs).cuda())
    assert torch.equal(x.cpu(), O.assemble_x(crops.numpy()))
    assert torch.equal(a.cpu(), O.window_audio(feats, idxs))
    u = model.forward_frames(crops.cuda(), feats.cuda(), torch.tensor(idxs).cuda())
    assert torch.equal(u, model.forward_uint8(x, a))


def test_forward_frames_staging_is_reused_safely():
    """forward_frames assembles (x, audio) in model-owned buffers (stable CUDA-graph keys).  Back-to-back calls with
    different inputs -- on one stream and on two streams without any synchronisation between them -- must each give the
    result of a fresh model: the second call may only overwrite the staging once the first forward has consumed it."""
    model, _ = make_model("R1", seed=6)
    fresh, _ = make_model("R1", seed=6)
    rs = np.random.RandomState(12)
    feats = torch.from_numpy(rs.randn(50, 2, 1024).astype(np.float32)).cuda()
    crops = [torch.from_numpy(rs.randint(0, 256, size=(6, 160, 160, 3), dtype=np.uint8)).cuda() for _ in range(4)]
    idx = [torch.arange(6 * i, 6 * i + 6, dtype=torch.int32).cuda() for i in range(4)]
    want = [fresh.forward_uint8(*fresh.prepare_inputs(crops[i], feats, idx[i])) for i in range(4)]
    torch.cuda.synchronize()
    got = [model.forward_frames(crops[0], feats, idx[0]), model.forward_frames(crops[1], feats, idx[1])]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for st, i in ((s1, 2), (s2, 3)):
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            got.append(model.forward_frames(crops[i], feats, idx[i]))
    torch.cuda.synchronize()
    for i in range(4):
        assert torch.equal(got[i], want[i]), i
    assert len(model._frame_io) == 1                  # one staging pair for the one batch size


def test_host_pipeline_frames_mode():
    """frames mode of HostPipeline == prepare_inputs + forward_uint8 called directly."""
    from calipsync_b200 import HostPipeline
    model, _ = make_model("R1")
    rs = np.random.RandomState(8)
    feats = torch.from_numpy(rs.randn(60, 2, 1024).astype(np.float32))
    pipe = HostPipeline(model, 4, frames=True)
    pipe.set_features(feats.pin_memory())
    batches = [(torch.from_numpy(rs.randint(0, 256, size=(n, 160, 160, 3), dtype=np.uint8)),
                torch.arange(n, dtype=torch.int32) + 4 * i) for i, n in enumerate((4, 4, 4, 2))]
    outs = [torch.empty(c.shape[0], 160, 160, 3, dtype=torch.uint8).pin_memory() for c, _ in batches]
    for (c, idx), o in zip(batches, outs):
        pipe.submit(c.pin_memory(), idx.pin_memory(), o)
    pipe.flush()
    fg = feats.cuda()
    for (c, idx), o in zip(batches, outs):
        x, a = model.prepare_inputs(c.cuda(), fg, idx.cuda())
        assert torch.equal(o, model.forward_uint8(x, a).cpu())


@pytest.mark.parametrize("hybrid", ["1", "0"])
@pytest.mark.parametrize("batch", [25, 70])
def test_two_lane_split_is_bit_exact(batch, hybrid, monkeypatch):
    """Batches >= 24 run as two half-batches on two streams: the low-resolution middle only (hybrid, the
    default: inc..down2 and up2..output run once for the whole batch) or the whole forward (CASYNC_HYBRID=0).  Frames
    are independent, so the result must equal the single-stream forward bit for bit -- odd batch sizes (unequal halves)
    included -- and the caller's stream must see the joined result without an explicit synchronisation.  With the
    hybrid split the lanes work on slices of ONE workspace layout, so the stage tensors are comparable too."""
    monkeypatch.setenv("CASYNC_HYBRID", hybrid)
    x, a = O.make_inputs(batch, 9)
    xs, as_ = x.cuda(), a.cuda()
    split, _ = make_model("R1", seed=3)
    out = split(xs, as_)
    doubled = out * 2.0                            # consumer on the caller's stream, enqueued right behind the forward
    n1 = split.launches_per_forward(8)
    if hybrid == "0":
        assert split.launches_per_forward(batch) == 2 * n1
    else:
        assert n1 < split.launches_per_forward(batch) < 2 * n1
    stages = {n: split.stage(n, batch).clone() for n in ("x3", "kx", "up1", "up3")} if hybrid == "1" else {}
    monkeypatch.setenv("CASYNC_SPLIT", "0")
    single, _ = make_model("R1", seed=3)
    ref = single(xs, as_)
    assert single.launches_per_forward(batch) == n1
    assert torch.equal(out, ref)
    assert torch.equal(doubled, ref * 2.0)
    for n, t in stages.items():
        assert torch.equal(single.stage(n, batch), t), n
    assert torch.equal(split.forward_uint8(xs, as_), single.forward_uint8(xs, as_))


@pytest.mark.parametrize("batch", [3, 40])
def test_cuda_graph_replay_matches_eager(batch, monkeypatch):
    """The second forward with the same tensors is captured into a CUDA graph and later calls replay it (one
    cudaGraphLaunch instead of 70-140 launches + fork / join events).  Replays must see NEW input values in the same
    buffers and give exactly the eager result; other tensors keep working next to the cached graph."""
    model, _ = make_model("R1", seed=4)
    x, a = O.make_inputs(batch, 21)
    xs, as_ = x.cuda(), a.cuda()
    out = torch.empty(batch, 3, 160, 160, device="cuda")
    run = lambda: model._run(xs, as_, out, 0).clone()   # noqa: E731  (fixed output buffer -> a stable key; flags 0 = bf16 path)
    first = run()                       # eager
    second = run()                      # captured + launched
    third = run()                       # replayed
    assert model.graph_replays() >= 2
    assert torch.equal(first, second) and torch.equal(first, third)
    x2, a2 = O.make_inputs(batch, 22)
    xs.copy_(x2.cuda())
    as_.copy_(a2.cuda())
    replayed = run()                    # same buffers, new contents
    monkeypatch.setenv("CASYNC_GRAPH", "0")
    eager_model, _ = make_model("R1", seed=4)
    ref = eager_model(xs, as_)
    assert eager_model.graph_replays() == 0
    assert torch.equal(replayed, ref)
    assert not torch.equal(replayed, first)
    assert torch.equal(model(xs, as_), ref)   # fresh output tensor: different key, still correct


@pytest.mark.parametrize("batch", [3, 30])
def test_depthwise_in_gemm_epilogue_is_bit_exact(batch, monkeypatch):
    """10x10 InvertedResidual blocks: the depthwise 3x3 runs in the epilogue of the first 1x1 conv (frame-aligned row
    tiles, hidden tile in shared memory) instead of its own launch.  Same operands, same order of operations: outputs and
    stage tensors must equal the three-launch path bit for bit."""
    monkeypatch.setenv("CASYNC_SPLIT", "0")
    monkeypatch.setenv("CASYNC_DWEPI", "1")      # opt-in (measured neutral at batch 64)
    x, a = O.make_inputs(batch, 31)
    fused, _ = make_model("R1", seed=5)
    out = fused(x.cuda(), a.cuda())
    stages = {n: fused.stage(n, batch).clone() for n in ("x5", "audio", "kx", "fuse")}
    monkeypatch.setenv("CASYNC_DWEPI", "0")
    plain, _ = make_model("R1", seed=5)
    ref = plain(x.cuda(), a.cuda())
    assert fused.launches_per_forward(batch) == plain.launches_per_forward(batch) - 7
    for n, t in stages.items():
        assert torch.equal(plain.stage(n, batch), t), n
    assert torch.equal(out, ref)
