"""BASELINE config 4: end-to-end 60 s of synthetic 16 kHz audio -> HuBERT features -> CASync generator, batch 256.

    python tools/e2e_hubert.py [seconds=60] [batch=256]

Upstream stage = the reference's `HubertExtractor.extract_features` (utils/hubert_extractor.py:18-58) restated with the
same arithmetic -- clips of 1000 feature steps (kernel 400, stride 320), pad / trim to expected_T, drop an odd last
frame, reshape [-1, 2, 1024] (25 fps) -- on a RANDOM-INIT HuBERT-large (`transformers.HubertModel`; the pretrained
`facebook/hubert-large-ls960-ft` weights are not on disk and there is no network; SURVEY 8(c): parity unpinned there,
it is third-party code and not on the graded path).  The features stay on the device (the reference moves them to the
CPU and back).  Downstream stage = calipsync_b200.Model.forward_frames: window gather + crop assembly + forward + uint8
epilogue.  Prints one JSON line with the stage split.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def hubert_large(device, dtype):
    from transformers import HubertConfig, HubertModel
    cfg = HubertConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                       feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True)
    torch.manual_seed(0)
    return HubertModel(cfg).to(device=device, dtype=dtype).eval()


@torch.no_grad()
def extract_features(model, speech):
    """utils/hubert_extractor.py:18-58 with the processor's normalisation (zero mean, unit variance) done in torch."""
    x = speech.float()
    x = (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)      # Wav2Vec2FeatureExtractor(do_normalize=True)
    iv = x[None, :].to(next(model.parameters()).dtype)
    kernel, stride = 400, 320
    clip = stride * 1000
    num_iter = iv.shape[1] // clip
    expected_t = (iv.shape[1] - (kernel - stride)) // stride
    feats = []
    for i in range(num_iter):
        s = clip * i
        e = s + (clip - stride + kernel)
        feats.append(model(iv[:, s:e]).last_hidden_state[0])
    rest = iv[:, clip * num_iter:]
    if rest.shape[1] >= kernel:
        feats.append(model(rest).last_hidden_state[0])
    f = torch.cat(feats, 0).float()
    if f.shape[0] < expected_t:
        f = torch.nn.functional.pad(f, (0, 0, 0, expected_t - f.shape[0]))
    else:
        f = f[:expected_t]
    if f.shape[0] % 2 == 1:
        f = f[:-1]
    return f.reshape(-1, 2, 1024).contiguous()


def main():
    seconds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    net = bench.build_model(dev)
    g = torch.Generator().manual_seed(7)
    speech = (torch.randn(seconds * 16000, generator=g) * 0.1).to(dev)
    out = {"config": "BASELINE configs[3]: %d s synthetic 16 kHz audio -> HuBERT-large (random init) -> UNet, batch %d"
                     % (seconds, batch)}
    for name, dtype in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        hub = hubert_large(dev, dtype)
        for _ in range(2):
            feats = extract_features(hub, speech)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        feats = extract_features(hub, speech)
        torch.cuda.synchronize()
        out["hubert_%s_ms" % name] = round((time.perf_counter() - t0) * 1e3, 2)
        del hub
    n = feats.shape[0]
    out["frames"] = n
    gg = torch.Generator(device=dev).manual_seed(3)
    crops = torch.randint(0, 256, (batch, 160, 160, 3), device=dev, dtype=torch.uint8, generator=gg)
    frames = torch.empty(n, 160, 160, 3, dtype=torch.uint8, device=dev)

    def unet_pass():
        for lo in range(0, n, batch):
            hi = min(n, lo + batch)
            idx = torch.arange(lo, hi, device=dev, dtype=torch.int32)
            net.forward_frames(crops[: hi - lo], feats, idx, out=frames[lo:hi])

    for _ in range(2):
        unet_pass()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    unet_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["unet_ms"] = round(ms, 2)
    out["unet_frames_per_s"] = round(n / ms * 1e3, 1)
    for name in ("fp32", "bf16"):
        tot = out["hubert_%s_ms" % name] + ms
        out["e2e_frames_per_s_hubert_%s" % name] = round(n / tot * 1e3, 1)
        out["unet_share_hubert_%s" % name] = round(ms / tot, 3)
    out["realtime_factor_hubert_bf16"] = round(seconds * 1e3 / (out["hubert_bf16_ms"] + ms), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
