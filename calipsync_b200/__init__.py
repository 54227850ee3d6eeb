"""calipsync_b200 -- B200-native (sm_100a) drop-in for the CASync generator forward pass.

Hot path only: ``Model.forward`` of the reference's ``module/unet.py`` (== ``image_infer_v1/models/unet.py``),
implemented as hand-written CUDA kernels behind a C ABI (``include/casync_b200.h``).  There is no CPU or
PyTorch fallback: without the built extension, or on a device that is not compute capability 10.x, the
forward raises.
"""
from .unet import Model  # noqa: F401
from .sharding import frame_shard, gather_chunked, shard_sizes, synthesize_clip  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from .blend import blend_paste  # noqa: F401

__all__ = ["Model", "HostPipeline", "frame_shard", "shard_sizes", "gather_chunked", "synthesize_clip", "blend_paste"]
