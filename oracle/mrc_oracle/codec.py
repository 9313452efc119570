"""Per-block encode/decode.  Follows /root/reference/codecThem.py: Decode :30-63, JointDecode :65-134,
calculateHuffmanGain :136-203, Encode :205-231, EncodeNoHuff :234-260, JointEncode :262-278,
EncodeSingleChannel :281-354, JointEncodeChannels :359-574.

codingParams is the reference's attribute bag (audiofile.py:51-53): a, b, nScaleBits, nMantSizeBits, sfBands,
targetBitsPerSample, sampleRate, nChannels, blkswBitA, blkswBitB and the mutable bitReservoir.

Mantissa representation: the reference returns either the int32 mantissa array (table 15) or a list of
strings "code" / "esccode/mantissa" (tables 0..3).  This restatement returns the same objects so that
pacfile.py can follow the reference's writer."""
import numpy as np

from .window import TransitionWindow
from .mdct import MDCT, IMDCT
from .quantize import ScaleFactor, vMantissa, vDequantize
from .ms_stereo import MSSwitchSFBands, ReconstructLR, OverallSMRs
from .psychoac import CalcSMRs
from .bitalloc import BitAlloc
from .tables import TABLES, NO_TABLE


def Decode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams):
    """:30-63  dequantise per band, undo the overall scale, IMDCT, window (no overlap-add here)."""
    cp = codingParams
    halfN = (cp.a + cp.b) // 2
    line = np.zeros(halfN, dtype=np.float64)
    i = 0
    for b in range(cp.sfBands.nBands):
        n = cp.sfBands.nLines[b]
        if bitAlloc[b]:
            line[i:i + n] = vDequantize(scaleFactor[b], mantissa[i:i + n], cp.nScaleBits, bitAlloc[b])
        i += n
    line /= 1. * (1 << overallScaleFactor)
    return TransitionWindow(IMDCT(line, cp.a, cp.b), cp.a, cp.b)


def JointDecode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, ms_switch):
    """:65-134  overallScaleFactor = [L, R, M, S]; channel 0 carries M|L, channel 1 carries S|R per band."""
    cp = codingParams
    lvl = [1. * (1 << s) for s in overallScaleFactor]
    halfN = (cp.a + cp.b) // 2
    l1 = np.zeros(halfN, dtype=np.float64)
    l2 = np.zeros(halfN, dtype=np.float64)
    i = 0
    for b in range(cp.sfBands.nBands):
        n = cp.sfBands.nLines[b]
        if bitAlloc[0][b]:
            l1[i:i + n] = vDequantize(scaleFactor[0][b], mantissa[0][i:i + n], cp.nScaleBits, bitAlloc[0][b])
            l1[i:i + n] /= lvl[2] if ms_switch[b] == 1 else lvl[0]
        if bitAlloc[1][b]:
            l2[i:i + n] = vDequantize(scaleFactor[1][b], mantissa[1][i:i + n], cp.nScaleBits, bitAlloc[1][b])
            l2[i:i + n] /= lvl[3] if ms_switch[b] == 1 else lvl[1]
        i += n
    left, right = ReconstructLR(l1, l2, cp.sfBands, ms_switch)
    return [TransitionWindow(IMDCT(left, cp.a, cp.b), cp.a, cp.b),
            TransitionWindow(IMDCT(right, cp.a, cp.b), cp.a, cp.b)]


def calculateHuffmanGain(mantissa, bitAlloc, codingParams):
    """:136-203.  raw = sum(ba*nLines); per table t (alphabetical order) cost = sum over transmitted mantissas
    of len_t[m] if m is a key of table t (this includes m == escape value: quirk Q4, the escape's own length
    only) else ba + len_t[escape]; the strict minimum below raw wins (ties -> lowest index), else table 15.
    The reference's early `break` once the running cost exceeds raw cannot change the choice (such a table can
    never be < the current minimum <= raw), so the full cost is summed here."""
    sf = codingParams.sfBands
    raw = 0
    per_line_ba = []
    for b in range(sf.nBands):
        if bitAlloc[b]:
            raw += int(bitAlloc[b]) * int(sf.nLines[b])
            per_line_ba += [int(bitAlloc[b])] * int(sf.nLines[b])
    m = [int(v) for v in mantissa]
    best, table = raw, NO_TABLE
    for t, T in enumerate(TABLES):
        cost = 0
        for v, ba in zip(m, per_line_ba):
            c = T.codes.get(v)
            cost += len(c) if c is not None else ba + len(T.escape_code)
        if cost < best:
            best, table = cost, t
    if table == NO_TABLE:
        codes = mantissa
    else:
        T = TABLES[table]
        codes = []
        for v in m:
            if v in T.codes and v != T.escape:
                codes.append(T.codes[v])
            else:
                codes.append(T.escape_code + "/" + str(v))      # escape code then the raw mantissa
    return table, codes, raw - best


def _budget_single(cp, halfN):
    """:299-308 (float arithmetic in the reference's order; Q12: the 4 huffTable bits are not budgeted)."""
    B = cp.targetBitsPerSample * halfN
    B -= cp.nScaleBits * (cp.sfBands.nBands + 1)
    B -= cp.nMantSizeBits * cp.sfBands.nBands
    B -= cp.blkswBitA
    B -= cp.blkswBitB
    B += cp.bitReservoir
    return B


def _budget_joint(cp, halfN):
    """:381-396"""
    B = cp.targetBitsPerSample * halfN
    B -= cp.nScaleBits * cp.sfBands.nBands
    B -= cp.nMantSizeBits * cp.sfBands.nBands
    B += B
    B -= cp.sfBands.nBands
    B -= cp.nScaleBits * 4
    B += cp.bitReservoir
    B -= cp.blkswBitA
    B -= cp.blkswBitB
    return B


def _max_mant_bits(cp):
    m = 1 << cp.nMantSizeBits
    return 16 if m > 16 else m


def _quantize_bands(lines_for_band, bitAlloc, cp):
    """:336-350 / :509-559  per band scale factor (computed even for zero-allocation bands, Q6) and the
    compacted mantissa array (zero-allocation bands omitted)."""
    sf = cp.sfBands
    scale = np.empty(sf.nBands, dtype=np.int32)
    nMant = sum(int(sf.nLines[b]) for b in range(sf.nBands) if bitAlloc[b])
    mant = np.empty(nMant, dtype=np.int32)
    i = 0
    for b in range(sf.nBands):
        lo, hi = sf.lowerLine[b], sf.upperLine[b] + 1
        x = lines_for_band(b)[lo:hi]
        scale[b] = ScaleFactor(np.max(np.abs(x)), cp.nScaleBits, bitAlloc[b])
        if bitAlloc[b]:
            n = sf.nLines[b]
            mant[i:i + n] = vMantissa(x, scale[b], cp.nScaleBits, bitAlloc[b])
            i += n
    return scale, mant


def EncodeSingleChannel(data, codingParams):
    """:281-354"""
    cp = codingParams
    halfN = (cp.a + cp.b) // 2
    B = _budget_single(cp, halfN)
    lines = MDCT(TransitionWindow(data, cp.a, cp.b), cp.a, cp.b)[:halfN]
    overall = ScaleFactor(np.max(np.abs(lines)), cp.nScaleBits)
    lines *= (1 << overall)
    SMRs = CalcSMRs(data, lines, overall, cp.sampleRate, cp.sfBands)
    bitAlloc, left = BitAlloc(B, _max_mant_bits(cp), cp.sfBands.nBands, cp.sfBands.nLines, SMRs)
    bitAlloc = bitAlloc.astype(int)
    cp.bitReservoir = int(left)
    scale, mant = _quantize_bands(lambda b: lines, bitAlloc, cp)
    cp._tap = dict(lines=[lines], smr=[SMRs], budget=B)
    return scale, bitAlloc, mant, overall


def JointEncodeChannels(dataLeft, dataRight, codingParams):
    """:359-574.  Mid/Side are formed in the time domain; the ms decision uses the unscaled L/R lines; the
    getMaskedThreshold / StereoMaskingFactor block at :465-474 is dead code (Q7) and skipped."""
    cp = codingParams
    sf = cp.sfBands
    dataMid = (dataLeft + dataRight) / 2.0
    dataSide = (dataLeft - dataRight) / 2.0
    halfN = (cp.a + cp.b) // 2
    B = _budget_joint(cp, halfN)
    sig = [dataLeft, dataRight, dataMid, dataSide]
    lines = [MDCT(TransitionWindow(x, cp.a, cp.b), cp.a, cp.b)[:halfN] for x in sig]
    ms_switch = MSSwitchSFBands(lines[0], lines[1], sf)
    overall = []
    for L in lines:
        s = ScaleFactor(np.max(np.abs(L)), cp.nScaleBits)
        L *= (1 << s)
        overall.append(s)
    smr = [CalcSMRs(sig[c], lines[c], overall[c], cp.sampleRate, sf) for c in range(4)]
    SMR1, SMR2 = OverallSMRs(smr[0], smr[1], smr[2], smr[3], sf, ms_switch)
    nLinesPass = np.append(sf.nLines, sf.nLines)
    SMRsPass = np.append(SMR1, SMR2)
    bitAlloc, left = BitAlloc(B, _max_mant_bits(cp), 2 * sf.nBands, nLinesPass, SMRsPass)
    bitAlloc = bitAlloc.astype(int)
    ba1, ba2 = bitAlloc[0:sf.nBands], bitAlloc[sf.nBands:]
    cp.bitReservoir = int(left)
    s1, m1 = _quantize_bands(lambda b: lines[2] if ms_switch[b] == 1 else lines[0], ba1, cp)
    s2, m2 = _quantize_bands(lambda b: lines[3] if ms_switch[b] == 1 else lines[1], ba2, cp)
    cp._tap = dict(lines=lines, smr=smr, smr_sel=[np.array(SMR1), np.array(SMR2)], budget=B)
    return [s1, s2], [ba1, ba2], [m1, m2], overall, ms_switch


def Encode(data, codingParams):
    """:205-231  channel iCh+1's budget sees channel iCh's Huffman savings through bitReservoir."""
    S, A, M, O, H = [], [], [], [], []
    taps = []
    for iCh in range(codingParams.nChannels):
        s, b, m, o = EncodeSingleChannel(data[iCh], codingParams)
        taps.append(codingParams._tap)
        t, codes, saved = calculateHuffmanGain(m, b, codingParams)
        codingParams.bitReservoir += saved
        S.append(s); A.append(b); M.append(codes); O.append(o); H.append(t)
    codingParams._tap = taps
    return S, A, M, O, H


def EncodeNoHuff(data, codingParams):
    """:234-260"""
    S, A, M, O, H = [], [], [], [], []
    for iCh in range(codingParams.nChannels):
        s, b, m, o = EncodeSingleChannel(data[iCh], codingParams)
        S.append(s); A.append(b); M.append(m); O.append(o); H.append(NO_TABLE)
    return S, A, M, O, H


def JointEncode(data, codingParams):
    """:262-278"""
    S, A, M, O, ms = JointEncodeChannels(data[0], data[1], codingParams)
    newM, H = [], []
    for iCh in range(codingParams.nChannels):
        t, codes, saved = calculateHuffmanGain(M[iCh], A[iCh], codingParams)
        codingParams.bitReservoir += saved
        H.append(t)
        newM.append(codes)
    return S, A, newM, O, ms, H
