"""Per-stage entry points of the C ABI, each fed with ORACLE inputs so errors cannot hide behind (or be blamed
on) upstream stages.  Compared with the oracle's sub-functions (module/unet.py sub-modules)."""
import ctypes

import pytest
import torch

from calipsync_b200 import Model, _lib
from oracle import casync_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1.2e-2     # rel-L2 of one block in bf16 (inputs rounded to bf16 + bf16 hidden tensors)


@pytest.fixture(scope="module")
def ctx():
    sd = O.make_state_dict(0, "R1")
    m = Model(6, "hubert")
    m.load_state_dict(sd)
    m = m.to("cuda:0").eval()
    x, a = O.make_inputs(2, 0)
    m(x.cuda(), a.cuda())                      # creates the plan
    _, st = O.forward(sd, x, a, return_stages=True)
    lib = _lib.load()
    scratch = torch.empty(lib.casync_stage_scratch_bytes(m._plan[0], 2), dtype=torch.uint8, device="cuda")
    return dict(sd=sd, m=m, x=x, a=a, st=st, lib=lib, scratch=scratch, plan=m._plan[0])


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw(t, b, h, c):
    return t.float().cpu().reshape(b, h, h, c).permute(0, 3, 1, 2)


_IR_STANDALONE = [i for i in range(1, 26) if i not in (18, 20, 22, 24)]   # 18..24 even: decoder blocks, test_up_first


@pytest.fixture(scope="module")
def ctx64(ctx):
    """Same model, scratch sized for BASELINE configs[1]'s batch (64 frames)."""
    scratch = torch.empty(ctx["lib"].casync_stage_scratch_bytes(ctx["plan"], 64), dtype=torch.uint8, device="cuda")
    return dict(ctx, scratch=scratch)


def _run_ir(c, idx, batch):
    d = _lib.ir_table()[idx]
    g = torch.Generator().manual_seed(idx)
    xin = (torch.randn(batch, d["cin"], d["h_in"], d["h_in"], generator=g) * 0.3).bfloat16().float()
    ref = O.inverted_residual(c["sd"], d["name"], xin, d["stride"], bool(d["residual"]))
    ho = d["h_in"] // d["stride"]
    out = torch.empty(batch * ho * ho, d["cout"], dtype=torch.bfloat16, device="cuda")
    xin_d = nhwc(xin)          # keep device inputs alive while the kernels run
    rc = c["lib"].casync_ir_block(c["plan"], idx, xin_d.data_ptr(), out.data_ptr(), c["scratch"].data_ptr(), batch,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "casync_ir_block")
    torch.cuda.synchronize()
    err = O.rel_l2(nchw(out, batch, ho, d["cout"]), ref)
    assert err < TOL, (d["name"], batch, err)


@pytest.mark.parametrize("idx", _IR_STANDALONE)
def test_inverted_residual_block(ctx, idx):
    """Every InvertedResidual instance that takes one input tensor (21 of 26; inc is covered by stage x1, the four
    decoder-first blocks by test_up_first), at batch 2."""
    _run_ir(ctx, idx, 2)


@pytest.mark.parametrize("idx", _IR_STANDALONE)
def test_inverted_residual_block_batch64(ctx64, idx):
    """The same blocks at BASELINE configs[1]'s batch: every CTA of the persistent kernels has several tiles / strips,
    frames straddle row tiles (M = 64 * H * W is not a multiple of 128 at 10x10)."""
    _run_ir(ctx64, idx, 64)


@pytest.mark.parametrize("level,batch", [(1, 2), (2, 2), (3, 2), (4, 2), (3, 64), (4, 64)])
def test_up_first(ctx, ctx64, level, batch):
    """up<level>.conv.double_conv.0 alone (IR indices 18, 20, 22, 24): bilinear x2 (align_corners) + concat + block.
    Levels 3 and 4 (the two hottest kernels of the step) also at batch 64."""
    import torch.nn.functional as F
    c = ctx64 if batch > 2 else ctx
    idx = 18 + 2 * (level - 1)
    d = _lib.ir_table()[idx]
    h = d["h_in"]
    g = torch.Generator().manual_seed(100 + level)
    low = (torch.randn(batch, d["cin"] // 2, h // 2, h // 2, generator=g) * 0.3).bfloat16().float()
    skip = (torch.randn(batch, d["cin"] // 2, h, h, generator=g) * 0.3).bfloat16().float()
    cat = torch.cat([F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=True), skip], dim=1)
    ref = O.inverted_residual(c["sd"], d["name"], cat, 1, False)
    out = torch.empty(batch * h * h, d["cout"], dtype=torch.bfloat16, device="cuda")
    low_d, skip_d = nhwc(low), nhwc(skip)
    rc = c["lib"].casync_up_first(c["plan"], level, low_d.data_ptr(), skip_d.data_ptr(), out.data_ptr(),
                                  c["scratch"].data_ptr(), batch, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "casync_up_first")
    torch.cuda.synchronize()
    err = O.rel_l2(nchw(out, batch, h, d["cout"]), ref)
    assert err < 2e-2, (d["name"], batch, err)


def test_audio_cnn(ctx):
    out = torch.empty(200, 512, dtype=torch.bfloat16, device="cuda")
    a_d = ctx["a"].cuda()
    rc = ctx["lib"].casync_audio_cnn(ctx["plan"], a_d.data_ptr(), out.data_ptr(), ctx["scratch"].data_ptr(), 2,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "casync_audio_cnn")
    torch.cuda.synchronize()
    assert O.rel_l2(nchw(out, 2, 10, 512), ctx["st"]["audio"]) < 2e-2


def test_fusion_attention(ctx):
    st = ctx["st"]
    kx = torch.empty(200, 1024, dtype=torch.bfloat16, device="cuda")
    x5_d, au_d = nhwc(st["x5"]), nhwc(st["audio"])
    rc = ctx["lib"].casync_fusion_attention(ctx["plan"], x5_d.data_ptr(), au_d.data_ptr(),
                                            kx.data_ptr(), ctx["scratch"].data_ptr(), 2,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "casync_fusion_attention")
    torch.cuda.synchronize()
    assert O.rel_l2(nchw(kx, 2, 10, 1024), st["kx"]) < 2e-2


@pytest.mark.parametrize("level", [1, 2, 3, 4])
def test_up_block(ctx, level):
    st = ctx["st"]
    low = st[("fuse", "up1", "up2", "up3")[level - 1]]
    skip = st[("x4", "x3", "x2", "x1")[level - 1]]
    ref = O.up_block(ctx["sd"], "up%d" % level, low.bfloat16().float(), skip.bfloat16().float())
    h, c = ref.shape[2], ref.shape[1]
    out = torch.empty(2 * h * h, c, dtype=torch.bfloat16, device="cuda")
    low_d, skip_d = nhwc(low), nhwc(skip)
    rc = ctx["lib"].casync_up_block(ctx["plan"], level, low_d.data_ptr(), skip_d.data_ptr(), out.data_ptr(),
                                    ctx["scratch"].data_ptr(), 2,
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "casync_up_block")
    torch.cuda.synchronize()
    assert O.rel_l2(nchw(out, 2, h, c), ref) < 2e-2


def test_bad_arguments_return_codes(ctx):
    lib = ctx["lib"]
    assert lib.casync_ir_block(ctx["plan"], 0, 1, 1, 1, 2, None) == -1       # inc is not runnable standalone
    assert lib.casync_ir_block(ctx["plan"], 99, 1, 1, 1, 2, None) == -1
    assert b"ir index" in lib.casync_last_error()
    assert lib.casync_forward(ctx["plan"], None, None, None, None, 1, 0, None) == -1
    assert lib.casync_forward(ctx["plan"], 256, 256, 256, 256, 1, _lib.F_FP32, None) == -4
