"""mrcaudiocodec_b200 -- B200-native (sm_100a) encode/decode hot path of the MRC perceptual audio codec.

Layout: csrc/ (CUDA kernels + C ABI, built into libmrc.so), _lib.py (ctypes binding), tables.py (host-side
constant tables), codec.py (batch API), codec_gpu.py (the reference's per-block seam: Encode / JointEncode /
Decode / JointDecode), pacfile.py (.pac container helpers), dist.py (multi-GPU sharding), synth.py (test signals).
There is no CPU implementation of the codec in this package."""
from .codec import Codec  # noqa: F401

__all__ = ["Codec"]
