"""vae_equalizer_b200 -- B200-native (sm_100a) training hot path of the VAE blind equalizer.

Python/PyTorch is the host: device memory, streams, torch.distributed.  All arithmetic of the hot
path runs in hand-written CUDA kernels behind the C ABI of libvaeq.so (include/vaeq.h).  There is
no CPU fallback: importing the operator modules without the built library, or calling them with
CPU tensors, raises.
"""
from ._lib import VaeqError, load, LIB_PATH  # noqa: F401

__all__ = ["VaeqError", "load", "LIB_PATH"]
