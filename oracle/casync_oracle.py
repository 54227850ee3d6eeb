"""CPU oracle for the CASync generator forward pass  --  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The shipped path (``calipsync_b200``) never imports ``oracle``
and fails loudly when its CUDA extension is missing.

What it is: a functional fp32 restatement (torch CPU ops on a plain ``dict`` of
tensors, no ``nn.Module``) of ``Model.forward`` in the reference's
``module/unet.py:314-345`` (identical copy: ``image_infer_v1/models/unet.py:281-312``).
Every function cites the reference lines it follows.

Parity pinning: the reference ships no tests, golden vectors or checkpoints for this
path ("parity unpinned" by the reference's own tests).  The oracle is
therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build
container by importing ``/root/reference/module/unet.py`` (script:
``tests/golden/make_golden.py``; fixtures: ``tests/golden/*.npz``), and -- whenever
``/root/reference`` is present -- live against the imported reference class
(``tests/test_oracle_vs_reference.py``).

Weights/inputs are produced by seeded generators defined here (``make_state_dict``,
``make_inputs``) so any machine can rebuild exactly the tensors the fixtures were
made from.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

CH = (32, 64, 128, 256, 512)   # module/unet.py:277
LEAKY_SLOPE = 0.01             # nn.LeakyReLU() default, module/unet.py:20,169,231,261,312
BN_EPS = 1e-5                  # nn.BatchNorm default
N_BLOCKS = 4                   # module/unet.py:274


# --------------------------------------------------------------------------------------
# state_dict spec (names, shapes, order) -- module/unet.py:273-312 registration order
# --------------------------------------------------------------------------------------
def _bn_entries(prefix, c):
    return [(prefix + ".weight", (c,), "bn_w", 0), (prefix + ".bias", (c,), "bn_b", 0),
            (prefix + ".running_mean", (c,), "bn_mean", 0), (prefix + ".running_var", (c,), "bn_var", 0),
            (prefix + ".num_batches_tracked", (), "bn_nbt", 0)]


def _ir_entries(prefix, inp, oup, expand=2):
    """InvertedResidual parameters, module/unet.py:16-34."""
    hid = inp * expand
    e = [(prefix + ".conv.0.weight", (hid, inp, 1, 1), "w", inp)]
    e += _bn_entries(prefix + ".conv.1", hid)
    e += [(prefix + ".conv.3.weight", (hid, 1, 3, 3), "w", 9)]
    e += _bn_entries(prefix + ".conv.4", hid)
    e += [(prefix + ".conv.6.weight", (oup, hid, 1, 1), "w", hid)]
    e += _bn_entries(prefix + ".conv.7", oup)
    return e


def _double_entries(prefix, cin, cout):
    """DoubleConvDW, module/unet.py:47-52."""
    return _ir_entries(prefix + ".double_conv.0", cin, cout) + _ir_entries(prefix + ".double_conv.1", cout, cout)


def _conv_entries(prefix, cout, cin, k):
    fan = cin * k * k
    return [(prefix + ".weight", (cout, cin, k, k), "w", fan), (prefix + ".bias", (cout,), "b", fan)]


def state_spec():
    """Ordered list of (name, shape, kind, fan_in) for Model(6, 'hubert', n_blocks=4)."""
    e = []
    a = "audio_model"                                          # module/unet.py:157-175
    e += _ir_entries(a + ".conv1", 32, CH[1]) + _ir_entries(a + ".conv2", CH[1], CH[2])
    e += _conv_entries(a + ".conv3", CH[3], CH[2], 3) + _bn_entries(a + ".bn3", CH[3])
    e += _ir_entries(a + ".conv4", CH[3], CH[3])
    e += _conv_entries(a + ".conv5", CH[4], CH[3], 3) + _bn_entries(a + ".bn5", CH[4])
    e += _ir_entries(a + ".conv6", CH[4], CH[4]) + _ir_entries(a + ".conv7", CH[4], CH[4])
    e += _bn_entries(a + ".bn7", CH[4])
    e += _double_entries("fuse_conv.0", CH[4] * 2, CH[4]) + _double_entries("fuse_conv.1", CH[4], CH[3])  # :286-289
    e += _ir_entries("inc.inconv.0", 6, CH[0])                 # :290
    for i in range(4):                                          # :291-294
        e += _double_entries("down%d.maxpool_conv.0" % (i + 1), CH[i], CH[i + 1])
    for i, (cin, cout) in enumerate([(CH[4], CH[3] // 2), (CH[3], CH[2] // 2), (CH[2], CH[1] // 2), (CH[1], CH[0])]):
        e += _double_entries("up%d.conv" % (i + 1), cin, cout)  # :296-299
    e += [("outc.conv.weight", (3, CH[0], 1, 1), "w", CH[0]), ("outc.conv.bias", (3,), "b", CH[0])]
    e += _bn_entries("outc_bn", 3)
    c2 = CH[4] * 2
    e += [("mlp_fusion.fc1.weight", (c2, c2), "w", c2), ("mlp_fusion.fc1.bias", (c2,), "b", c2)]
    e += _bn_entries("mlp_fusion.bn1", c2)
    e += [("mlp_fusion.fc2.weight", (c2, c2), "w", c2), ("mlp_fusion.fc2.bias", (c2,), "b", c2)]
    e += _bn_entries("mlp_fusion.bn2", c2)
    for i in range(N_BLOCKS):                                   # :306-308, :252-261, :198-205
        p = "attention_blocks.%d" % i
        e += [(p + ".cross_attention.gamma", (1,), "gamma", 0)]
        e += _conv_entries(p + ".cross_attention.query_conv", CH[4] // 8, CH[4], 1)
        e += _conv_entries(p + ".cross_attention.key_conv", CH[4] // 8, CH[4], 1)
        e += _conv_entries(p + ".cross_attention.value_conv", CH[4], CH[4], 1)
        e += _conv_entries(p + ".attention_adjust_p_1", CH[4], c2, 1)
        e += _conv_entries(p + ".attention_adjust_b_1", c2, CH[4], 1)
        e += _bn_entries(p + ".bn", c2)
    e += _bn_entries("bn_kx", c2) + _bn_entries("bn_tx", c2)
    return e


def _gen(seed, name):
    h = int.from_bytes(hashlib.sha256(("%d:%s" % (seed, name)).encode()).digest()[:7], "little")
    return torch.Generator().manual_seed(h)


def make_state_dict(seed=0, regime="R1"):
    """Seeded weights.  R0: torch-default-like init (kaiming-uniform(a=sqrt5) == U(-1/sqrt(fan_in), ..),
    BN identity, gamma 0 -- module/unet.py:205).  R1: R0 + perturbed BN statistics/affine and gamma=0.5
    (the primary parity regime; exercises BN folding and the attention path)."""
    assert regime in ("R0", "R1")
    sd = OrderedDict()
    for name, shape, kind, fan in state_spec():
        g = _gen(seed, name)
        if kind in ("w", "b"):
            bound = 1.0 / math.sqrt(fan)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "gamma":
            t = torch.full(shape, 0.5 if regime == "R1" else 0.0)
        elif kind == "bn_nbt":
            t = torch.zeros((), dtype=torch.int64)
        elif regime == "R0":
            t = torch.ones(shape) if kind in ("bn_w", "bn_var") else torch.zeros(shape)
        elif kind in ("bn_w", "bn_var"):
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        else:  # bn_b, bn_mean
            t = 0.1 * torch.randn(shape, generator=g)
        sd[name] = t
    return sd


def make_inputs(batch, seed=0, frame_offset=0):
    """Synthetic inputs with the reference's layout.  x = cat([face, masked face]) in [0,1] with the
    cv2.rectangle((5,5,150,145)) mask zeroing rows 5..149 x cols 5..154 of channels 3..5
    (image_infer_v1/tools/frame_synthesizer/infer_api.py:238-245, dataset/dataset.py:98);
    audio = a [16,2,1024] HuBERT window reshaped to [32,32,32] (dataset/dataset.py:172-176).
    Frame i depends only on (seed, frame_offset+i) so any rank can build its own shard."""
    xs, auds = [], []
    for i in range(batch):
        g = _gen(seed, "frame%d" % (frame_offset + i))
        x = torch.rand(6, 160, 160, generator=g)
        x[3:6, 5:150, 5:155] = 0.0
        xs.append(x)
        auds.append(torch.randn(32, 32, 32, generator=g))
    return torch.stack(xs), torch.stack(auds)


def window_audio(features, indices):
    """HuBERT windowing, image_infer_v1/tools/frame_synthesizer/infer_api.py:99-145 (and
    dataset/dataset.py:39-56): rows idx-8..idx+8 of [T,2,1024], zero-padded at clip ends,
    reshaped to [32,32,32].

    Exact, including the reference's quirk (:126-129): the padding is built as ``zeros_like(auds[:pad])``, i.e. it can
    never be longer than the rows already collected; when that truncates it the window has fewer than 16 rows and the
    reference falls back to an ALL-ZERO feature (:133-145).  For frame indices inside a clip of >= 16 feature frames
    this never triggers; it does for indices past the clip end."""
    T = features.shape[0]
    out = torch.zeros(len(indices), 16, 2, 1024, dtype=features.dtype)
    for n, idx in enumerate(indices):
        lo, hi = idx - 8, idx + 8
        pad_l, pad_r = max(0, -lo), max(0, hi - T)
        s_lo, s_hi = max(lo, 0), min(hi, T)
        n_valid = s_hi - s_lo
        if n_valid > 0 and pad_l <= n_valid and pad_r <= n_valid + pad_l:
            out[n, s_lo - lo:s_hi - lo] = features[s_lo:s_hi]
    return out.reshape(len(indices), 32, 32, 32)


# --------------------------------------------------------------------------------------
# the forward pass
# --------------------------------------------------------------------------------------
def assemble_x(crops_u8):
    """Caller-side image assembly, image_infer_v1/tools/frame_synthesizer/infer_api.py:238-245 (twin:
    dataset/dataset.py:129-132): crops_u8 uint8 [B,160,160,3] (= crop_img[4:164, 4:164]) -> fp32 [B,6,160,160] =
    cat([crop, masked crop]) / 255 with cv2.rectangle(img, (5, 5, 150, 145), (0,0,0), -1) = rows 5..149 x cols 5..154
    zeroed.  numpy float32 arithmetic, exactly as the reference."""
    import numpy as np
    crops = np.asarray(crops_u8)
    out = []
    for img in crops:
        masked = img.copy()
        masked[5:150, 5:155] = 0
        real = img.transpose(2, 0, 1).astype(np.float32) / 255.0
        masked = masked.transpose(2, 0, 1).astype(np.float32) / 255.0
        out.append(np.concatenate([real, masked]))
    return torch.from_numpy(np.stack(out))


def blend_paste(img, crop_resized, face_mask_u8, rect, soft_mask=None):
    """The paste-back blend of FrameSynthesizer.process_batch, image_infer_v1/tools/frame_synthesizer/infer_api.py:313-346,
    for one frame (numpy, float64 like the reference): img uint8 [H,W,3] (a modified copy is returned), crop_resized
    uint8 [h,w,3] = the re-sized crop with the prediction pasted in, face_mask_u8 uint8 [h,w] = the dilated polygon
    (`final_face_mask`), rect = (ymin, ymax, xmin, xmax), soft_mask float32 [h,w] = the re-sized per-frame mask or None."""
    import numpy as np
    ymin, ymax, xmin, xmax = rect
    m3 = np.repeat((face_mask_u8 / 255.0)[..., np.newaxis], 3, axis=2)                      # :313-314
    if soft_mask is not None:
        s3 = np.repeat(soft_mask[..., np.newaxis], 3, axis=2)                               # :336
        inverted = 1.0 - s3                                                                 # :339
        m3 = m3 * (1.0 - inverted)                                                          # :342
    out = img.copy()
    out[ymin:ymax, xmin:xmax] = (crop_resized * m3) + (img[ymin:ymax, xmin:xmax] * (1.0 - m3))   # :345 / :348, :350
    return out


def _bn(sd, p, x):
    """eval-mode BatchNorm (running statistics), eps=1e-5."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _leaky(x):
    return F.leaky_relu(x, LEAKY_SLOPE)


def inverted_residual(sd, p, x, stride, res):
    """module/unet.py:8-40: pw expand x2 -> BN -> leaky -> dw3x3(stride,pad1) -> BN -> leaky -> pw project
    -> BN -> leaky; residual added AFTER the last activation (:38)."""
    h = _leaky(_bn(sd, p + ".conv.1", F.conv2d(x, sd[p + ".conv.0.weight"])))
    h = F.conv2d(h, sd[p + ".conv.3.weight"], None, stride, 1, 1, h.shape[1])
    h = _leaky(_bn(sd, p + ".conv.4", h))
    h = _leaky(_bn(sd, p + ".conv.7", F.conv2d(h, sd[p + ".conv.6.weight"])))
    return x + h if res else h


def double_conv(sd, p, x, stride):
    """module/unet.py:43-55."""
    x = inverted_residual(sd, p + ".double_conv.0", x, stride, False)
    return inverted_residual(sd, p + ".double_conv.1", x, 1, True)


def audio_encoder(sd, a):
    """AudioConvHubert.forward, module/unet.py:177-194 (conv5 has padding 3: 16 -> 10)."""
    p = "audio_model"
    a = inverted_residual(sd, p + ".conv1", a, 1, False)
    a = inverted_residual(sd, p + ".conv2", a, 1, False)
    a = _leaky(_bn(sd, p + ".bn3", F.conv2d(a, sd[p + ".conv3.weight"], sd[p + ".conv3.bias"], 2, 1)))
    a = inverted_residual(sd, p + ".conv4", a, 1, True)
    a = _leaky(_bn(sd, p + ".bn5", F.conv2d(a, sd[p + ".conv5.weight"], sd[p + ".conv5.bias"], 2, 3)))
    a = inverted_residual(sd, p + ".conv6", a, 1, True)
    a = inverted_residual(sd, p + ".conv7", a, 1, True)
    return _leaky(_bn(sd, p + ".bn7", a))


def mlp_fusion(sd, x, y):
    """MLPFusion.forward, module/unet.py:233-249: per-position fc over cat([visual, audio]) channels;
    BatchNorm1d acts on the channel dim."""
    B, C, H, W = x.shape
    f = torch.cat([x.reshape(B, C, -1).permute(0, 2, 1), y.reshape(B, C, -1).permute(0, 2, 1)], dim=-1)
    f = F.linear(f, sd["mlp_fusion.fc1.weight"], sd["mlp_fusion.fc1.bias"])
    f = _leaky(_bn(sd, "mlp_fusion.bn1", f.permute(0, 2, 1))).permute(0, 2, 1)
    f = F.linear(f, sd["mlp_fusion.fc2.weight"], sd["mlp_fusion.fc2.bias"])
    f = _bn(sd, "mlp_fusion.bn2", f.permute(0, 2, 1))
    return f.reshape(B, -1, H, W)


def cross_attention(sd, p, x, y):
    """CrossAttention.forward, module/unet.py:207-218: queries = visual tokens, keys/values = audio
    tokens, softmax over keys with NO 1/sqrt(d) scale, out = V . attn^T, gamma*out + x."""
    B, C, H, W = x.shape
    q = F.conv2d(x, sd[p + ".query_conv.weight"], sd[p + ".query_conv.bias"]).reshape(B, -1, H * W).permute(0, 2, 1)
    k = F.conv2d(y, sd[p + ".key_conv.weight"], sd[p + ".key_conv.bias"]).reshape(B, -1, H * W)
    attn = F.softmax(torch.bmm(q, k), dim=-1)
    v = F.conv2d(y, sd[p + ".value_conv.weight"], sd[p + ".value_conv.bias"]).reshape(B, -1, H * W)
    out = torch.bmm(v, attn.permute(0, 2, 1)).reshape(B, C, H, W)
    return sd[p + ".gamma"] * out + x


def attention_block(sd, p, x, audio, tx):
    """AttentionBlock.forward, module/unet.py:263-270 (ox + tx happens BEFORE the BN)."""
    ox = F.conv2d(x, sd[p + ".attention_adjust_p_1.weight"], sd[p + ".attention_adjust_p_1.bias"])
    ox = cross_attention(sd, p + ".cross_attention", ox, audio)
    ox = F.conv2d(ox, sd[p + ".attention_adjust_b_1.weight"], sd[p + ".attention_adjust_b_1.bias"])
    return _leaky(_bn(sd, p + ".bn", ox + tx))


def up_block(sd, p, x1, x2):
    """Up.forward, module/unet.py:90-97: bilinear x2 align_corners=True, pad to the skip size, cat([up, skip])."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return double_conv(sd, p + ".conv", torch.cat([x1, x2], dim=1), 1)


STAGE_NAMES = ("x1", "x2", "x3", "x4", "x5", "audio", "tx", "ox0", "ox1", "ox2", "ox3", "kx",
               "fuse", "up1", "up2", "up3", "up4", "logits", "out")


@torch.no_grad()
def forward(sd, x, audio_feat, return_stages=False):
    """Model.forward, module/unet.py:314-345.  x: [B,6,160,160] fp32, audio_feat: [B,32,32,32] fp32
    -> [B,3,160,160] fp32 in (0,1).  With return_stages also returns the NCHW stage activations."""
    st = OrderedDict()
    x1 = inverted_residual(sd, "inc.inconv.0", x, 1, False)                 # :315
    x2 = double_conv(sd, "down1.maxpool_conv.0", x1, 2)                    # :316
    x3 = double_conv(sd, "down2.maxpool_conv.0", x2, 2)
    x4 = double_conv(sd, "down3.maxpool_conv.0", x3, 2)
    x5 = double_conv(sd, "down4.maxpool_conv.0", x4, 2)                    # :319
    a = audio_encoder(sd, audio_feat)                                       # :321
    tx = _bn(sd, "bn_tx", torch.cat([x5, a], dim=1) + mlp_fusion(sd, x5, a))  # :323-326
    st.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, audio=a, tx=tx)
    ox = kx = tx
    for i in range(N_BLOCKS):                                               # :331-333
        ox = attention_block(sd, "attention_blocks.%d" % i, ox, a, tx)
        kx = ox + kx
        st["ox%d" % i] = ox
    kx = _leaky(_bn(sd, "bn_kx", kx))                                       # :335-336
    f = double_conv(sd, "fuse_conv.1", double_conv(sd, "fuse_conv.0", kx, 1), 1)   # :337
    u1 = up_block(sd, "up1", f, x4)                                         # :338-341
    u2 = up_block(sd, "up2", u1, x3)
    u3 = up_block(sd, "up3", u2, x2)
    u4 = up_block(sd, "up4", u3, x1)
    logits = _bn(sd, "outc_bn", F.conv2d(u4, sd["outc.conv.weight"], sd["outc.conv.bias"]))   # :342-343
    out = torch.sigmoid(logits)                                             # :344
    st.update(kx=kx, fuse=f, up1=u1, up2=u2, up3=u3, up4=u4, logits=logits, out=out)
    return (out, st) if return_stages else out


# --------------------------------------------------------------------------------------
# parity metrics (BASELINE.json north_star tolerances)
# --------------------------------------------------------------------------------------
def max_abs_255(a, b):
    return float((a.double() - b.double()).abs().max() * 255.0)


def psnr_db(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 (b = reference)."""
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
