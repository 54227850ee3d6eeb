"""Weight packer: reference ``state_dict`` (582 entries, module/unet.py:273-312) -> one blob in the layout the
sm_100a kernels consume.

* eval-mode BatchNorm (eps 1e-5) is folded in float64 into the preceding conv / linear: scale into the
  weight rows, shift into a bias vector (module/unet.py:18,28,32; :163,:168,:174; :228,:230; :260; :301,:310-311);
* every dense weight becomes bf16 "UMMA tiles": for k-block kb (64 input channels) and output row n the
  128 bytes are stored as the SWIZZLE_128B shared-memory image (16-byte chunk c at position c ^ (n & 7)), so a
  kernel fetches a whole [BN x 64] B-operand tile with one bulk-async copy and feeds it to tcgen05.mma;
* depthwise weights are fp32 [tap][channel]; tiny layers (inc, outc) are fp32 structs passed as kernel
  parameters.

The library publishes the schema (entry names/sizes: ``casync_weight_entry``); this module fills it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

BN_EPS = 1e-5


def _bn_fold(sd, p):
    """y = s*x + t of an eval-mode BatchNorm."""
    s = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + BN_EPS)
    t = sd[p + ".bias"].double() - sd[p + ".running_mean"].double() * s
    return s, t


def pack_gemm_weight(w: torch.Tensor) -> np.ndarray:
    """[N, K] float -> uint8 [(K/64) * N * 128]: bf16, k-block major, rows of 64 with the 128B swizzle."""
    n, k = w.shape
    kb = (k + 63) // 64
    wp = torch.zeros(n, kb * 64, dtype=torch.float64)
    wp[:, :k] = w.double()
    wb = wp.to(torch.bfloat16).view(n, kb, 8, 8)                       # [n, kb, chunk, 8]
    src_chunk = torch.arange(8)[None, :] ^ (torch.arange(n)[:, None] & 7)  # physical p holds logical p ^ (n&7)
    idx = src_chunk[:, None, :, None].expand(n, kb, 8, 8)
    sw = torch.gather(wb, 2, idx).permute(1, 0, 2, 3).contiguous()     # [kb, n, chunk, 8]
    return sw.view(torch.uint8).numpy().reshape(-1).copy()


def unpack_gemm_weight(raw: np.ndarray, n: int, k: int) -> torch.Tensor:
    """Inverse of pack_gemm_weight (tests only): uint8 blob -> fp32 [N, K] of the bf16 values."""
    kb = (k + 63) // 64
    sw = torch.from_numpy(np.ascontiguousarray(raw)).view(torch.bfloat16).view(kb, n, 8, 8).permute(1, 0, 2, 3)
    src_chunk = torch.arange(8)[None, :] ^ (torch.arange(n)[:, None] & 7)
    idx = src_chunk[:, None, :, None].expand(n, kb, 8, 8)
    return torch.gather(sw, 2, idx).reshape(n, kb * 64)[:, :k].float()


def _f32(t) -> np.ndarray:
    return t.to(torch.float32).contiguous().numpy().view(np.uint8).reshape(-1).copy()


def _ir_parts(sd, p):
    s1, t1 = _bn_fold(sd, p + ".conv.1")
    s2, t2 = _bn_fold(sd, p + ".conv.4")
    s3, t3 = _bn_fold(sd, p + ".conv.7")
    w1 = sd[p + ".conv.0.weight"].double().flatten(1) * s1[:, None]          # [hid, cin]
    wd = (sd[p + ".conv.3.weight"].double().flatten(1) * s2[:, None]).t()    # [9, hid]
    w2 = sd[p + ".conv.6.weight"].double().flatten(1) * s3[:, None]          # [cout, hid]
    return w1, t1, wd, t2, w2, t3


def build_entries(sd):
    """name -> uint8 numpy array for every schema entry."""
    sd = {k: v.detach().cpu() for k, v in sd.items()}
    out = {}
    ir_cache = {}

    def ir_parts(prefix):
        if prefix not in ir_cache:
            ir_cache[prefix] = _ir_parts(sd, prefix)
        return ir_cache[prefix]

    for name, _ in _lib.weight_schema():
        prefix, part = name.split("|")
        if part == "wdp":
            # depthwise taps + folded-BN bias as bf16, [hid/8][10][8]: the 64-channel slice a depthwise work item of
            # the layer-program kernel handles is 1280 contiguous bytes (one bulk copy next to its slab of hidden rows)
            _, _, wd, bd, _, _ = ir_parts(prefix)
            hid = wd.shape[1]
            t = torch.cat([wd, bd[None, :]], 0).view(10, hid // 8, 8).permute(1, 0, 2).contiguous()
            val = t.float().to(torch.bfloat16).view(torch.uint8).numpy().reshape(-1).copy()   # same two roundings as the fp32 taps
        elif part in ("w1", "b1", "wd", "bd", "w2", "b2"):
            w1, b1, wd, bd, w2, b2 = ir_parts(prefix)
            val = {"w1": lambda: pack_gemm_weight(w1), "b1": lambda: _f32(b1), "wd": lambda: _f32(wd),
                   "bd": lambda: _f32(bd), "w2": lambda: pack_gemm_weight(w2), "b2": lambda: _f32(b2)}[part]()
        elif part == "inc":
            w1, b1, wd, bd, w2, b2 = ir_parts(prefix)
            val = _f32(torch.cat([w1.reshape(-1), b1, wd.reshape(-1), bd, w2.reshape(-1), b2]))
        elif part == "w2t":   # pw2 of `inc` for the tensor core: [32, 12] -> one k-block, K zero-padded to 64
            val = pack_gemm_weight(ir_parts(prefix)[4])
        elif prefix in ("audio_model.conv3", "audio_model.conv5"):
            bn = prefix.replace("conv", "bn")
            s, t = _bn_fold(sd, bn)
            if part == "w":   # K ordered (tap, cin) to match the implicit-GEMM A producer
                w = sd[prefix + ".weight"].double().permute(0, 2, 3, 1).flatten(1) * s[:, None]
                val = pack_gemm_weight(w)
            else:
                val = _f32(s * sd[prefix + ".bias"].double() + t)
        elif prefix == "audio_model.bn7":
            s, t = _bn_fold(sd, prefix)
            val = _f32(s if part == "s" else t)
        elif prefix == "mlp_fusion.fc1":
            s, t = _bn_fold(sd, "mlp_fusion.bn1")
            val = pack_gemm_weight(sd[prefix + ".weight"].double() * s[:, None]) if part == "w" else \
                _f32(s * sd[prefix + ".bias"].double() + t)
        elif prefix == "mlp_fusion.fc2":
            # tx = bn_tx(cat + bn2(fc2(h)))  (module/unet.py:244-246, 323-326)
            s2, t2 = _bn_fold(sd, "mlp_fusion.bn2")
            st, tt = _bn_fold(sd, "bn_tx")
            if part == "w":
                val = pack_gemm_weight(sd[prefix + ".weight"].double() * (st * s2)[:, None])
            elif part == "b":
                val = _f32(st * (s2 * sd[prefix + ".bias"].double() + t2) + tt)
            else:
                val = _f32(st)
        elif prefix == "attention_blocks":
            # rows ordered [k_0 .. k_3 (4 x 64) | v_0 | v_1 | v_2 | v_3 (4 x 512)]: the key block is stored row-major,
            # the value blocks transposed per frame (GemmArgs::vt) for the tensor-core attention kernel
            ca = "attention_blocks.%d.cross_attention."
            ws = [sd[(ca % j) + "key_conv.weight"].double().flatten(1) for j in range(4)] + \
                 [sd[(ca % j) + "value_conv.weight"].double().flatten(1) for j in range(4)]
            bs = [sd[(ca % j) + "key_conv.bias"].double() for j in range(4)] + \
                 [sd[(ca % j) + "value_conv.bias"].double() for j in range(4)]
            if part == "kv_w":
                val = pack_gemm_weight(torch.cat(ws, 0))
            elif part == "kv_b":
                val = _f32(torch.cat(bs, 0))
            else:
                val = _f32(torch.cat([sd["attention_blocks.%d.cross_attention.gamma" % j].double().reshape(1)
                                      for j in range(4)]))
        elif prefix.startswith("attention_blocks."):
            if part.startswith("p1q"):
                # [attention_adjust_p_1 ; query_conv o attention_adjust_p_1]  (module/unet.py:264, 209): one GEMM
                # emits p_1(x) and q = Wq(Wp x + bp) + bq
                wp = sd[prefix + ".attention_adjust_p_1.weight"].double().flatten(1)
                bp = sd[prefix + ".attention_adjust_p_1.bias"].double()
                wq = sd[prefix + ".cross_attention.query_conv.weight"].double().flatten(1)
                bq = sd[prefix + ".cross_attention.query_conv.bias"].double()
                val = pack_gemm_weight(torch.cat([wp, wq @ wp], 0)) if part == "p1q_w" else \
                    _f32(torch.cat([bp, wq @ bp + bq], 0))
            else:  # ox = leaky(bn(b_1(.) + tx))  (module/unet.py:266-269): the BN scale also multiplies tx
                m = prefix + ".attention_adjust_b_1"
                s, t = _bn_fold(sd, prefix + ".bn")
                if part == "b1_w":
                    val = pack_gemm_weight(sd[m + ".weight"].double().flatten(1) * s[:, None])
                elif part == "b1_b":
                    val = _f32(s * sd[m + ".bias"].double() + t)
                else:
                    val = _f32(s)
        elif prefix == "bn_kx":
            s, t = _bn_fold(sd, prefix)
            val = _f32(s if part == "s" else t)
        elif prefix == "outc":
            s, t = _bn_fold(sd, "outc_bn")
            w = sd["outc.conv.weight"].double().flatten(1) * s[:, None]
            b = s * sd["outc.conv.bias"].double() + t
            val = _f32(torch.cat([w.reshape(-1), b, torch.zeros(1, dtype=torch.float64)]))
        else:
            raise KeyError("packer: unknown schema entry %r" % name)
        out[name] = val
    return out


def pack(sd):
    """-> (blob uint8 tensor [bytes] on CPU, offsets int64 numpy [n_entries])."""
    schema = _lib.weight_schema()
    entries = build_entries(sd)
    offsets, cur = [], 0
    for name, nbytes in schema:
        if entries[name].size != nbytes:
            raise ValueError("packer: entry %s has %d bytes, library expects %d" % (name, entries[name].size, nbytes))
        offsets.append(cur)
        cur = (cur + nbytes + 255) & ~255
    blob = np.zeros(cur, dtype=np.uint8)
    for (name, nbytes), off in zip(schema, offsets):
        blob[off:off + nbytes] = entries[name]
    return torch.from_numpy(blob), np.asarray(offsets, dtype=np.int64)
