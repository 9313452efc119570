"""16-bit PCM <-> signed fraction.  Follows /root/reference/pcmfile.py: ReadDataBlock :87-101 (with
quantize.vDequantizeUniform :90-111) and WriteDataBlock :164-174 (with quantize.vQuantizeUniform :61-87).
RIFF header parsing (:34-66, :141-153) is host plumbing and out of scope."""
import numpy as np
from .quantize import vDequantizeUniform, vQuantizeUniform


def pcm_to_fraction(codes_int16):
    """sign/magnitude split, dequantise the magnitude with 16 bits, restore the sign.  |c| = 32768 is read as
    'negative, magnitude 0' by the dequantiser, so -32768 -> 0.0 (quirk Q1)."""
    c = np.asarray(codes_int16).astype(np.int64)
    neg = np.signbit(c)
    c = np.where(neg, -c, c)
    x = vDequantizeUniform(c, 16)
    x[neg] *= -1.
    return x


def fraction_to_pcm(x):
    """|x|>=1 -> 32767 else trunc((65535|x|+1)/2), then the sign."""
    t = np.array(x, dtype=np.float64, copy=True)
    neg = np.signbit(t)
    t[neg] *= -1.
    q = vQuantizeUniform(t, 16).astype(np.int16)
    q[neg] *= -1
    return q
