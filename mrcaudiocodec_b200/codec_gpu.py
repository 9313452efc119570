"""codec_gpu -- the reference's per-block seam, served by libmrc.so.

Drop-in for the module the reference imports as `codec` (pacfileThem.py:108 `import codecThem as codec`): the same
function names, argument meaning and tuple shapes as codecThem.py

    Encode(data, codingParams)          -> (scaleFactor[ch], bitAlloc[ch], mantissa[ch], overallScaleFactor[ch], huffTable[ch])   :205-231
    EncodeNoHuff(data, codingParams)    -> same, tables forced to 15                                                               :234-260
    JointEncode(data, codingParams)     -> (scaleFactor[2], bitAlloc[2], mantissa[2], overallScale[4]=[L,R,M,S], ms_switch, huffTable[2]) :262-278
    Decode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams)            -> windowed IMDCT output, one channel    :30-63
    JointDecode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, ms_switch) -> [left, right]                     :65-134

so a reference-style PACFile can be pointed at the GPU with `pacfileThem.codec = codec_gpu`.
`data` is a list of per-channel float64 arrays of length a+b; codingParams is the reference's attribute bag
(reads a, b, nScaleBits, nMantSizeBits, targetBitsPerSample, sampleRate, nChannels; reads AND writes bitReservoir).
mantissa[ch] is the compacted int32 array (table 15) or the list of "code" / "esccode/mantissa" strings
(tables 0..3), exactly what calculateHuffmanGain returns (codecThem.py:178-203).

One or two channels (nChannels = 1: Encode / EncodeNoHuff / Decode, the reference's loops over
codingParams.nChannels, codecThem.py:216; the whole-file batch API is stereo).  Block switching: codingParams.a / codingParams.b may each be nMDCTLines or
128, as the reference's `__main__` loop sets them (pacfileThem.py:1193-1210); the band table follows a + b like
pacfileThem.py:808-816 (25 bands for two long halves, else the 9-band short table), nMDCTLines must be 1024 then.
Every call is one round trip to the GPU: this layer is for parity and integration, the batch API in codec.py is
the fast path."""
import json
import os

import numpy as np

from .codec import Codec

_here = os.path.dirname(os.path.abspath(__file__))
_ctx_cache = {}
_tables = None


def _huff():
    global _tables
    if _tables is None:
        with open(os.path.join(_here, "huffman_tables.json")) as fh:
            d = json.load(fh)
        _tables = [({int(k): v for k, v in t["codes"].items()}, int(t["escape"])) for t in d["tables"]]
    return _tables


def _codec_for(cp, precision=None):
    for v in (cp.a, cp.b):
        if v != cp.nMDCTLines and v != 128:
            raise NotImplementedError("codec_gpu serves window halves of nMDCTLines or 128 samples")
    if getattr(cp, "nChannels", 2) not in (1, 2):
        raise NotImplementedError("codec_gpu serves one- and two-channel streams")
    precision = precision or getattr(cp, "precision", "fp64")
    switched = not (cp.a == cp.b == cp.nMDCTLines)
    key = (int(cp.sampleRate), int(cp.nMDCTLines), int(cp.nScaleBits), int(cp.nMantSizeBits),
           float(getattr(cp, "targetBitsPerSample", 0.0)), precision, int(getattr(cp, "device", 0)))
    c = _ctx_cache.get(key)
    if c is None or (switched and len(c.block_tables) == 1):
        if c is not None:
            c.close()
        c = Codec(sample_rate=key[0], n_mdct_lines=key[1], n_scale_bits=key[2], n_mant_size_bits=key[3],
                  target_bits_per_sample=key[4], precision=precision, device=key[6], switch_tables=switched)
        _ctx_cache[key] = c
    return c


def _compact(c, r, ch, a, b, as_codes=True):
    """line-aligned mantissas -> what the reference returns for channel ch."""
    _, nbands, n, lo = c.geometry(a, b)
    ba = r["bitAlloc"][ch]
    parts = [r["mantissa"][ch][lo[k]:lo[k] + n[k]] for k in range(nbands) if ba[k]]
    m = np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    t = int(r["huffTable"][ch])
    if t == 15 or not as_codes:
        return m
    codes, esc = _huff()[t]
    out = []
    for v in m.tolist():
        if v in codes and v != esc:
            out.append(codes[v])
        else:
            out.append(codes[esc] + "/" + str(v))
    return out


def _encode(data, codingParams, joint, no_huff=False):
    c = _codec_for(codingParams)
    nch = int(getattr(codingParams, "nChannels", 2))
    if nch == 1:
        if joint:
            raise ValueError("JointEncode needs two channels (codecThem.py:363-364)")
        x = np.asarray(data[0], dtype=np.float64)[None, :]
    else:
        x = np.stack([np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64)])
    r, res = c.encode_block(x, (1 if joint else 0) | (2 if no_huff else 0) | (4 if nch == 1 else 0),
                            int(codingParams.bitReservoir), a=codingParams.a, b=codingParams.b)
    codingParams.bitReservoir = res
    S = [r["scaleFactor"][ch].astype(np.int32) for ch in range(nch)]
    A = [r["bitAlloc"][ch].astype(int) for ch in range(nch)]
    M = [_compact(c, r, ch, codingParams.a, codingParams.b) for ch in range(nch)]
    H = [int(r["huffTable"][ch]) for ch in range(nch)]
    return c, r, S, A, M, H


def Encode(data, codingParams):
    c, r, S, A, M, H = _encode(data, codingParams, joint=False)
    return (S, A, M, [int(r["overallScale"][ch]) for ch in range(len(S))], H)


def EncodeNoHuff(data, codingParams):
    c, r, S, A, M, H = _encode(data, codingParams, joint=False, no_huff=True)
    return (S, A, M, [int(r["overallScale"][ch]) for ch in range(len(S))], H)


def JointEncode(data, codingParams):
    c, r, S, A, M, H = _encode(data, codingParams, joint=True)
    return (S, A, M, [int(v) for v in r["overallScale"]], [int(v) for v in r["ms_switch"]], H)


def Decode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams):
    c = _codec_for(codingParams)
    a, b = codingParams.a, codingParams.b
    nl, nbands, _, _ = c.geometry(a, b)
    z = np.zeros(nbands, np.int32)
    m = np.asarray(mantissa, dtype=np.int32)
    y = c.decode_block(False, [scaleFactor, z], [bitAlloc, z], np.stack([m[:nl], np.zeros(nl, np.int32)]),
                       [overallScaleFactor, 0], a=a, b=b)
    return y[0]


def JointDecode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, ms_switch):
    c = _codec_for(codingParams)
    a, b = codingParams.a, codingParams.b
    nl = c.geometry(a, b)[0]
    m = np.stack([np.asarray(mantissa[0], dtype=np.int32)[:nl], np.asarray(mantissa[1], dtype=np.int32)[:nl]])
    y = c.decode_block(True, scaleFactor, bitAlloc, m, overallScaleFactor, ms_switch, a=a, b=b)
    return [y[0], y[1]]
