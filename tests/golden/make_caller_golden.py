"""Golden fixture from the UNMODIFIED caller (run in the build container only; needs /root/reference).

    python tests/golden/make_caller_golden.py

Builds a tiny synthetic ``data_dir`` (JPEG frames + landmark files, the layout of step3_prepare_infer_data.py), a
checkpoint written with ``torch.save(state_dict)`` and runs the reference's own
``image_infer_v1/tools/frame_synthesizer/infer_api.py::FrameSynthesizer`` on CPU with the REFERENCE ``Model``:
``_load_batch_frames`` -> ``_get_audio_features`` -> ``process_batch`` (crop, mask, model call, uint8 truncation,
paste-back + polygon blend), nothing patched except a recorder around ``self.net``.  Recorded into
``caller_batch.npz``: the uint8 160x160 crops the caller fed to the model (its float input is exactly
``cat([crop, masked crop]) / 255``), the HuBERT windows, the reference model's truncated uint8 predictions and the
blended full frames.  tests/test_caller_contract.py replays the same scene here (CPU, reference present) and
tests/test_gpu_parity.py checks the CUDA path against the recorded predictions on the GPU box, where the reference does
not exist.
"""
import os
import sys
import tempfile

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, "/root/reference")

from oracle import casync_oracle as O  # noqa: E402

N_FRAMES, SIZE, T_FEAT = 3, 320, 12


def build_scene(root):
    """data_dir with N_FRAMES smooth random frames and 110-point landmark files around a face box."""
    rs = np.random.RandomState(2024)
    os.makedirs(os.path.join(root, "frames"))
    os.makedirs(os.path.join(root, "positions"))
    os.makedirs(os.path.join(root, "masks"))
    for i in range(N_FRAMES):
        img = cv2.GaussianBlur((rs.rand(SIZE, SIZE, 3) * 255).astype(np.uint8), (0, 0), 3)
        cv2.imwrite(os.path.join(root, "frames", "%06d.jpg" % i), img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        x0, y0, w = 70 + 3 * i, 60 + 2 * i, 170 + 4 * i            # face box: lms[1].x, lms[52].y, lms[31].x - lms[1].x
        lms = np.zeros((110, 2), dtype=np.float64)
        ang = np.linspace(0, 2 * np.pi, 33, endpoint=False)         # face contour polygon (points 0..32)
        lms[:33, 0] = x0 + w / 2 + 0.48 * w * np.cos(ang)
        lms[:33, 1] = y0 + w / 2 + 0.48 * w * np.sin(ang)
        lms[33:, 0] = x0 + rs.rand(77) * w
        lms[33:, 1] = y0 + rs.rand(77) * w
        lms[1] = (x0, y0 + w / 2)
        lms[31] = (x0 + w, y0 + w / 2)
        lms[52] = (x0 + w / 2, y0)
        np.savetxt(os.path.join(root, "positions", "%06d.txt" % i), lms)
    feats = rs.randn(T_FEAT, 2, 1024).astype(np.float32)
    return feats


class Recorder(torch.nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net, self.calls = net, []

    def forward(self, x, a):
        out = self.net(x, a)
        self.calls.append((x.clone(), a.clone(), out.clone()))
        return out


def run_caller(model_cls, device, root, feats, ckpt):
    import image_infer_v1.models.unet as ref_unet
    import image_infer_v1.tools.frame_synthesizer.infer_api as caller
    keep = caller.Model
    caller.Model = model_cls
    try:
        fs = caller.FrameSynthesizer(ckpt, root, device=device, batch_size=N_FRAMES)
    finally:
        caller.Model = keep
    del ref_unet
    fs.net = Recorder(fs.net)
    idx = list(range(N_FRAMES))
    images, lms, masks = fs._load_batch_frames(idx)
    hub = fs._get_audio_features(feats, [0, 5, T_FEAT - 1])            # clip start, interior, clip end (zero padded)
    results = fs.process_batch(images, lms, masks, hub)
    fs.executor.shutdown()
    return images, hub, results, fs.net.calls


def main():
    from image_infer_v1.models.unet import Model as RefModel
    with tempfile.TemporaryDirectory() as root:
        feats = build_scene(root)
        ckpt = os.path.join(root, "unet.pth")
        torch.save(O.make_state_dict(5, "R1"), ckpt)
        images, hub, results, calls = run_caller(RefModel, "cpu", root, feats, ckpt)
    assert len(calls) == 1
    x, a, out = calls[0]
    crops = (x[:, :3] * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()     # exact: x = u8 / 255
    assert torch.equal(O.assemble_x(crops.numpy()), x)
    pred_u8 = np.array(out.numpy().transpose(0, 2, 3, 1) * 255, dtype=np.uint8)           # the caller's own truncation
    np.savez_compressed(os.path.join(HERE, "caller_batch.npz"), crops=crops.numpy(), audio=a.numpy(), pred_u8=pred_u8)
    print("caller_batch.npz:", os.path.getsize(os.path.join(HERE, "caller_batch.npz")) // 1024, "KB;",
          "pred range", float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
