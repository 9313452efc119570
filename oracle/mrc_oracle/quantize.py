"""Quantisers.  Follows /root/reference/quantize.py: QuantizeUniform :12-38, vQuantizeUniform :61-87,
vDequantizeUniform :90-111, ScaleFactor :114-146, vMantissa :294-322, vDequantize :325-357.
The vector routines are restated in closed form on integers; tests/test_oracle_vs_reference.py checks
them value for value against the shimmed reference."""
import math
import numpy as np


def QuantizeUniform(aNum, nBits):
    """:12-38  sign-magnitude midtread; code = trunc(((2^nBits-1)*|x| + 1)/2), clipped at |x|>=1."""
    sign = 0 if aNum >= 0.0 else 1
    if abs(aNum) >= 1:
        code = pow(2, nBits - 1) - 1
    else:
        code = int(np.int64(((pow(2, nBits) - 1) * abs(aNum) + 1) / 2))
    return (sign << (nBits - 1)) + code


def vQuantizeUniform(aNumVec, nBits):
    """:61-87  returned as float64 like the reference (sign vector is float)."""
    nBits = int(nBits)
    x = np.asarray(aNumVec, dtype=np.float64)
    mag = np.absolute(x)
    t = np.divide(np.add(np.multiply(pow(2.0, nBits) - 1, mag), 1.0), 2.0)   # mul, add, div: three roundings
    code = np.where(mag < 1.0, np.int64(np.where(mag < 1.0, t, 0.0)), np.int64(pow(2.0, nBits - 1) - 1.0))
    code = np.where(mag == 0.0, np.int64(0), code)
    return np.add(np.where(x < 0.0, float(1 << (nBits - 1)), 0.0), code)


def vDequantizeUniform(aQuantizedNumVec, nBits):
    """:90-111  x = sign * mag * 2 / (2^nBits - 1)."""
    q = np.asarray(aQuantizedNumVec, dtype=np.float64)
    neg = q >= pow(2.0, nBits - 1)
    mag = np.where(neg, q - pow(2.0, nBits - 1), q)
    sign = np.where(neg, -1.0, 1.0)
    return np.divide(np.multiply(np.multiply(sign, mag), 2.0), pow(2, nBits) - 1.0)


def ScaleFactor(aNum, nScaleBits=3, nMantBits=5):
    """:114-146  leading zeros of the (2^nScaleBits-1+nMantBits)-bit magnitude code, capped at 2^nScaleBits-1.
    Uses math.log(code, 2) like the reference."""
    nBits = int(pow(2, nScaleBits) - 1 + nMantBits)
    quant = QuantizeUniform(aNum, nBits)
    magCode = quant - pow(2, nBits - 1) if quant >= pow(2, nBits - 1) else quant
    top = 0 if magCode == 0 else int(math.log(magCode, 2))
    lz = (nBits - 2) - top
    cap = pow(2, nScaleBits) - 1
    return lz if lz < cap else cap


def vMantissa(aNumVec, scale, nScaleBits=3, nMantBits=5):
    """:294-322  block floating point: sign bit at 2^(nMantBits-1) plus magnitude code >> (cap - scale)."""
    nBits = pow(2, nScaleBits) - 1 + nMantBits
    q = vQuantizeUniform(aNumVec, nBits)
    neg = q >= pow(2, nBits - 1)
    mag = np.where(neg, q - pow(2, nBits - 1), q).astype(np.uint64)
    cap = pow(2, nScaleBits) - 1
    if scale != cap:
        mag = np.right_shift(mag, np.uint64(cap - scale))
    return np.add(np.where(neg, float(pow(2, nMantBits - 1)), 0.0), mag)


def vDequantize(scale, mantissaVec, nScaleBits=3, nMantBits=5):
    """:325-357  inverse of vMantissa; adds the half-step 2^(shift-1) to non-zero magnitudes when shifted."""
    nBits = pow(2, nScaleBits) - 1 + nMantBits
    m = np.asarray(mantissaVec, dtype=np.float64)
    neg = m >= pow(2, nMantBits - 1)
    mag = np.where(neg, m - pow(2, nMantBits - 1), m)
    cap = pow(2, nScaleBits) - 1
    if scale == cap:
        code = mag
    else:
        shift = cap - scale
        code = np.left_shift(mag.astype(np.uint64), np.uint64(shift)).astype(np.float64)
        if shift > 0:
            code = code + np.where(mag > 0, float(pow(2, shift - 1)), 0.0)
    return vDequantizeUniform(np.add(code, np.where(neg, float(pow(2, nBits - 1)), 0.0)), nBits)
