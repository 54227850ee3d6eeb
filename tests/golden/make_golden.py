"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/module/unet.py::Model`` (never copied into this repo), loads the seeded
state_dict from ``oracle.casync_oracle.make_state_dict`` with ``strict=True`` and records, for weight
regimes R0 and R1, the reference's own fp32 CPU output for the seeded 2-frame input of
``make_inputs(2, seed=0)`` plus, per stage (forward hooks on the reference modules), a strided probe
of 256 values and the stage L2 norm.  The fixtures pin ``oracle/`` (tests/test_oracle_golden.py, CPU)
and the CUDA path (tests/test_gpu_parity.py) on machines where ``/root/reference`` does not exist.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

from module.unet import Model  # noqa: E402  (the reference)
from oracle import casync_oracle as O  # noqa: E402

HOOKS = {"x1": "inc", "x2": "down1", "x3": "down2", "x4": "down3", "x5": "down4", "audio": "audio_model",
         "tx": "bn_tx", "ox0": "attention_blocks.0", "ox1": "attention_blocks.1", "ox2": "attention_blocks.2",
         "ox3": "attention_blocks.3", "kx": "lru_kx", "fuse": "fuse_conv", "up1": "up1", "up2": "up2",
         "up3": "up3", "up4": "up4", "logits": "outc_bn"}


def probe(t):
    flat = t.reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, 256).long()
    return flat[idx].numpy().astype(np.float32)


def main():
    torch.manual_seed(0)
    for regime in ("R0", "R1"):
        sd = O.make_state_dict(0, regime)
        net = Model(6, "hubert").eval()
        net.load_state_dict(sd, strict=True)
        x, a = O.make_inputs(2, 0)
        got = {}
        mods = dict(net.named_modules())
        hs = [mods[m].register_forward_hook(lambda _m, _i, o, k=k: got.__setitem__(k, o.detach()))
              for k, m in HOOKS.items()]
        with torch.no_grad():
            out = net(x, a)
        for h in hs:
            h.remove()
        rec = {"out": out.numpy().astype(np.float32)}
        for k, v in got.items():
            rec["probe_" + k] = probe(v)
            rec["norm_" + k] = np.float64(v.double().norm().item())
        path = os.path.join(HERE, "unet_%s_seed0_b2.npz" % regime)
        np.savez_compressed(path, **rec)
        print(path, os.path.getsize(path) // 1024, "KiB", "out range", float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
