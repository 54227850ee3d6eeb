"""Packer: UMMA tile layout round trip and the BatchNorm-folding algebra, proven on CPU against the oracle."""
import numpy as np
import pytest
import torch

from calipsync_b200 import _lib, packer
from oracle import casync_oracle as O
from packed_emulation import PackedNet


@pytest.mark.parametrize("n,k", [(8, 64), (72, 100), (64, 32), (256, 1152), (32, 12)])
def test_gemm_weight_roundtrip(n, k):
    w = torch.randn(n, k)
    raw = packer.pack_gemm_weight(w)
    assert raw.dtype == np.uint8 and raw.size == ((k + 63) // 64) * n * 128
    assert torch.equal(packer.unpack_gemm_weight(raw, n, k), w.bfloat16().float())


def test_swizzle_positions():
    """16-byte chunk c of row n sits at chunk position c ^ (n & 7) of its 128-byte row (SWIZZLE_128B)."""
    w = torch.arange(16 * 64, dtype=torch.float32).reshape(16, 64) % 251
    raw = torch.from_numpy(packer.pack_gemm_weight(w)).view(torch.bfloat16).view(16, 8, 8).float()
    for n in (0, 1, 5, 7, 8, 13):
        for c in range(8):
            assert torch.equal(raw[n, c ^ (n & 7)], w[n, c * 8:(c + 1) * 8].bfloat16().float())


def test_blob_matches_library_schema():
    blob, offsets = packer.pack(O.make_state_dict(0, "R1"))
    schema = _lib.weight_schema()
    assert len(offsets) == len(schema) and all(o % 256 == 0 for o in offsets)
    assert blob.numel() >= offsets[-1] + schema[-1][1]
    assert 39e6 < blob.numel() < 41e6        # ~19.8 M parameters in bf16


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_folded_weights_reproduce_oracle(regime):
    sd = O.make_state_dict(0, regime)
    x, a = O.make_inputs(1, 0)
    ref, rst = O.forward(sd, x, a, return_stages=True)
    with torch.no_grad():
        out, st = PackedNet(sd, round_bf16=False).forward(x, a)   # only the weights are bf16
    for name in O.STAGE_NAMES:
        assert O.rel_l2(st[name], rst[name]) < 6e-3, name
    assert O.max_abs_255(out, ref) < 0.05
    with torch.no_grad():
        out, _ = PackedNet(sd, round_bf16=True).forward(x, a)      # + bf16 activations between kernels
    assert O.max_abs_255(out, ref) < 2.0 and O.psnr_db(out, ref) > 45.0


def test_packed_depthwise_taps_layout():
    """`wdp` = depthwise taps + folded-BN bias as bf16, [hid/8][10][8]: entry (chunk, t, e) is tap t (t = 9: the bias)
    of channel 8*chunk + e, rounded fp64 -> fp32 -> bf16 exactly like the fp32 taps the fused kernel converts on the
    device -- the stand-alone depthwise kernel, the weight-streaming fused blocks and the strip kernels read it."""
    sd = O.make_state_dict(0, "R1")
    entries = packer.build_entries(sd)
    for prefix, hid in (("down4.maxpool_conv.0.double_conv.1", 1024), ("audio_model.conv4", 512)):
        wdp = torch.from_numpy(entries[prefix + "|wdp"]).view(torch.bfloat16).view(hid // 8, 10, 8).float()
        wd = torch.from_numpy(entries[prefix + "|wd"]).view(torch.float32).view(9, hid)
        bd = torch.from_numpy(entries[prefix + "|bd"]).view(torch.float32)
        for t in range(9):
            assert torch.equal(wdp[:, t, :].reshape(-1), wd[t].bfloat16().float()), (prefix, t)
        assert torch.equal(wdp[:, 9, :].reshape(-1), bd.bfloat16().float()), prefix
