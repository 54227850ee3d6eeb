"""CPU emulation of the CUDA launch sequence over the PACKED weights (test helper, not product code).

Consumes exactly what the kernels consume -- the entries produced by ``calipsync_b200.packer.build_entries``
(folded BatchNorm, bf16 UMMA tiles, fp32 depthwise taps) -- and replays csrc/plan.cu's sequence with torch
fp32 ops, optionally rounding activations to bf16 at the points where the kernels store bf16.  Used to
(1) prove the BN-folding algebra of the packer against the oracle on CPU, and (2) predict the error budget
of the bf16 path.  Layout here is NCHW fp32 for convenience; only the arithmetic mirrors the kernels.
"""
import numpy as np
import torch
import torch.nn.functional as F

from calipsync_b200 import _lib, packer

LEAK = 0.01


def _f(raw):
    return torch.from_numpy(np.ascontiguousarray(raw)).view(torch.float32).clone()


class PackedNet:
    def __init__(self, sd, round_bf16=True):
        self.e = packer.build_entries(sd)
        self.rb = round_bf16
        self.ir = _lib.ir_table()

    def r(self, t):
        return t.bfloat16().float() if self.rb else t

    def gemm_w(self, name, n, k):
        return packer.unpack_gemm_weight(self.e[name], n, k)

    def pw(self, x, wname, bname, n, k):
        w = self.gemm_w(wname, n, k)
        return F.conv2d(x, w[:, :, None, None], _f(self.e[bname]))

    def ir_block(self, idx, x, post=None):
        d = self.ir[idx]
        p, cin, hid, cout = d["name"] + "|", d["cin"], 2 * d["cin"], d["cout"]
        if idx == 0:  # fused fp32 inc kernel: weights fp32, hidden never rounded
            v = _f(self.e[d["name"] + "|inc"])
            w1, b1, wd, bd, w2, b2 = torch.split(v, [72, 12, 108, 12, 384, 32])
            h = F.leaky_relu(F.conv2d(x, w1.view(12, 6, 1, 1), b1), LEAK)
            h = F.leaky_relu(F.conv2d(h, wd.view(9, 12).t().reshape(12, 1, 3, 3), bd, 1, 1, 1, 12), LEAK)
            return self.r(F.leaky_relu(F.conv2d(h, w2.view(32, 12, 1, 1), b2), LEAK))
        h = self.r(F.leaky_relu(self.pw(x, p + "w1", p + "b1", hid, cin), LEAK))
        wd = _f(self.e[p + "wd"]).view(9, hid).t().reshape(hid, 1, 3, 3)
        h = self.r(F.leaky_relu(F.conv2d(h, wd, _f(self.e[p + "bd"]), d["stride"], 1, 1, hid), LEAK))
        y = F.leaky_relu(self.pw(h, p + "w2", p + "b2", cout, hid), LEAK)
        if d["residual"]:
            y = y + x
        if post is not None:
            y = F.leaky_relu(post[0].view(1, -1, 1, 1) * y + post[1].view(1, -1, 1, 1), LEAK)
        return self.r(y)

    def dense3x3(self, x, pre, cin, cout, pad):
        w = self.gemm_w(pre + "|w", cout, 9 * cin).view(cout, 3, 3, cin).permute(0, 3, 1, 2)
        return self.r(F.leaky_relu(F.conv2d(x, w, _f(self.e[pre + "|b"]), 2, pad), LEAK))

    def forward(self, x, audio):
        st = {}
        x1 = self.ir_block(0, x)
        cur, skips = x1, [x1]
        for l in range(4):
            cur = self.ir_block(2 + 2 * l, self.ir_block(1 + 2 * l, cur))
            skips.append(cur)
        x1, x2, x3, x4, x5 = skips
        a = self.ir_block(10, self.ir_block(9, self.r(audio)))
        a = self.ir_block(11, self.dense3x3(a, "audio_model.conv3", 128, 256, 1))
        a = self.ir_block(12, self.dense3x3(a, "audio_model.conv5", 256, 512, 3))
        a = self.ir_block(13, a, post=(_f(self.e["audio_model.bn7|s"]), _f(self.e["audio_model.bn7|t"])))
        cat = torch.cat([x5, a], 1)
        h = self.r(F.leaky_relu(self.pw(cat, "mlp_fusion.fc1|w", "mlp_fusion.fc1|b", 1024, 1024), LEAK))
        tx = self.r(self.pw(h, "mlp_fusion.fc2|w", "mlp_fusion.fc2|b", 1024, 1024)
                    + _f(self.e["mlp_fusion.fc2|rs"]).view(1, -1, 1, 1) * cat)
        kv = self.r(self.pw(a, "attention_blocks|kv_w", "attention_blocks|kv_b", 2304, 512))
        gamma = _f(self.e["attention_blocks|gamma"])
        st.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, audio=a, tx=tx)
        B = x.shape[0]
        ox, oxs = tx, []
        for j in range(4):
            p = "attention_blocks.%d|" % j
            p1q = self.r(self.pw(ox, p + "p1q_w", p + "p1q_b", 576, 1024))
            p1, q = p1q[:, :512], p1q[:, 512:].reshape(B, 64, 100)
            k = kv[:, j * 64: (j + 1) * 64].reshape(B, 64, 100)
            v = kv[:, 256 + j * 512: 256 + (j + 1) * 512].reshape(B, 512, 100)
            attn = torch.softmax(torch.bmm(q.permute(0, 2, 1), k), -1)
            o = torch.bmm(v, attn.permute(0, 2, 1)).reshape(B, 512, 10, 10)
            att = self.r(gamma[j] * o + p1)
            ox = self.r(F.leaky_relu(self.pw(att, p + "b1_w", p + "b1_b", 1024, 512)
                                     + _f(self.e[p + "b1_rs"]).view(1, -1, 1, 1) * tx, LEAK))
            oxs.append(ox)
            st["ox%d" % j] = ox
        kx = tx + oxs[0] + oxs[1] + oxs[2] + oxs[3]
        kx = self.r(F.leaky_relu(_f(self.e["bn_kx|s"]).view(1, -1, 1, 1) * kx + _f(self.e["bn_kx|t"]).view(1, -1, 1, 1), LEAK))
        f = kx
        for i in range(14, 18):
            f = self.ir_block(i, f)
        u, ups = f, []
        for lvl, skip in enumerate((x4, x3, x2, x1)):
            up = self.r(F.interpolate(u, scale_factor=2, mode="bilinear", align_corners=True))
            u = self.ir_block(19 + 2 * lvl, self.ir_block(18 + 2 * lvl, torch.cat([up, skip], 1)))
            ups.append(u)
        v = _f(self.e["outc|outc"])
        logits = F.conv2d(u, v[:96].view(3, 32, 1, 1), v[96:99])
        out = torch.sigmoid(logits)
        st.update(kx=kx, fuse=f, up1=ups[0], up2=ups[1], up3=ups[2], up4=ups[3], logits=logits, out=out)
        return out, st
