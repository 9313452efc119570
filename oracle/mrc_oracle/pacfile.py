"""The .pac container.  Follows /root/reference/pacfileThem.py (class PACFile): ReadFileHeader :130-158,
ReadDataBlock :161-319, JointReadDataBlock :321-585, WriteFileHeader :586-619, WriteDataBlock :622-790,
JointWriteDataBlock :793-972, Close :973-984.  Works on in-memory bytes instead of an open file.

Layout as the shipped code writes it:
  header  'PAC ' | <LHLLHH sampleRate nChannels numSamples nMDCTLines nScaleBits nMantSizeBits | <L nBands |
          <{nBands}H nLines
  per block, per channel:  <L nBytes | nBytes of MSB-first bits, zero padded:
     huffTable(4) blkswA(1) blkswB(1)
     non-joint: overallScale(nScaleBits)       joint, channel 0 only: overallScale L,R,M,S then nBands ms bits
     per band: (ba ? ba-1 : 0)(nMantSizeBits) scaleFactor(nScaleBits) then, if ba, nLines mantissas:
               ba raw bits, or the Huffman code; escape = escape code + ba raw bits
"""
import io
from struct import pack, unpack, calcsize

import numpy as np

from .bitpack import PackedBits, BYTESIZE
from .psychoac import ScaleFactorBands, AssignMDCTLinesFromFreqLimits, shortFreqLimits
from .tables import TABLES, NO_TABLE
from . import codec

TAG = b'PAC '
N_HUFF_BITS = 4


class CodingParams(object):
    """audiofile.py:51-53"""
    pass


def sfbands_for(cp):
    """pacfileThem.py:208-219 etc.: 25 critical bands for two long halves, else the 9-band short table."""
    half = (cp.a + cp.b) // 2
    if cp.a + cp.b == 2 * cp.nMDCTLines:
        return ScaleFactorBands(AssignMDCTLinesFromFreqLimits(half, cp.sampleRate))
    return ScaleFactorBands(AssignMDCTLinesFromFreqLimits(half, cp.sampleRate, shortFreqLimits))


class PACWriter(object):
    def __init__(self, codingParams):
        self.buf = io.BytesIO()
        self.chunks = []                    # (offset, nBytes) of every channel chunk, for tests
        self.WriteFileHeader(codingParams)

    def WriteFileHeader(self, cp):
        """:586-619  (numSamples quirk Q9: bumped by nMDCTLines only when already a multiple)."""
        self.buf.write(TAG)
        if not cp.numSamples % cp.nMDCTLines:
            cp.numSamples += (cp.nMDCTLines - cp.numSamples % cp.nMDCTLines)
        self.buf.write(pack('<LHLLHH', cp.sampleRate, cp.nChannels, cp.numSamples, cp.nMDCTLines,
                            cp.nScaleBits, cp.nMantSizeBits))
        sf = ScaleFactorBands(AssignMDCTLinesFromFreqLimits(cp.nMDCTLines, cp.sampleRate))
        cp.sfBands = sf
        self.buf.write(pack('<L', sf.nBands))
        self.buf.write(pack('<' + str(sf.nBands) + 'H', *(sf.nLines.tolist())))
        cp.priorBlock = [np.zeros(cp.nMDCTLines, dtype=np.float64) for _ in range(cp.nChannels)]

    # -- one channel chunk ---------------------------------------------------------------------
    def _chunk(self, cp, head_fields, bitAlloc, scaleFactor, mantissa, huffTable):
        """size computation :651-707 / :825-880 and packing :716-789 / :894-970."""
        sf = cp.sfBands
        nBits = N_HUFF_BITS + cp.blkswBitA + cp.blkswBitB + sum(w for _, w in head_fields)
        iMant = 0
        for b in range(sf.nBands):
            nBits += cp.nMantSizeBits + cp.nScaleBits
            if bitAlloc[b]:
                if huffTable == NO_TABLE:
                    nBits += int(bitAlloc[b]) * int(sf.nLines[b])
                else:
                    esc = TABLES[huffTable].escape_code
                    for _ in range(int(sf.nLines[b])):
                        code = str(mantissa[iMant]).split("/")[0]
                        nBits += len(code) + (int(bitAlloc[b]) if code == esc else 0)
                        iMant += 1
        nBytes = nBits // BYTESIZE if nBits % BYTESIZE == 0 else nBits // BYTESIZE + 1
        self.buf.write(pack("<L", int(nBytes)))
        pb = PackedBits()
        pb.Size(nBytes)
        pb.WriteBits(huffTable, N_HUFF_BITS)
        pb.WriteBits(1 - cp.a // cp.nMDCTLines, cp.blkswBitA)
        pb.WriteBits(1 - cp.b // cp.nMDCTLines, cp.blkswBitB)
        for v, w in head_fields:
            pb.WriteBits(v, w)
        iMant = 0
        for b in range(sf.nBands):
            ba = int(bitAlloc[b])
            pb.WriteBits(ba - 1 if ba else 0, cp.nMantSizeBits)
            pb.WriteBits(int(scaleFactor[b]), cp.nScaleBits)
            if ba:
                for _ in range(int(sf.nLines[b])):
                    if huffTable == NO_TABLE:
                        pb.WriteBits(int(mantissa[iMant]), ba)
                    else:
                        parts = str(mantissa[iMant]).split("/")
                        for ch in parts[0]:
                            pb.WriteBits(int(ch), 1)
                        if parts[0] == TABLES[huffTable].escape_code:
                            pb.WriteBits(int(parts[1]), ba)
                    iMant += 1
        self.chunks.append((self.buf.tell(), int(nBytes)))
        self.buf.write(pb.GetPackedData())

    def WriteDataBlock(self, data, cp):
        """:622-790  independent channels."""
        full = [np.concatenate((cp.priorBlock[c], data[c])) for c in range(cp.nChannels)]
        cp.priorBlock = data
        cp.sfBands = sfbands_for(cp)
        S, A, M, O, H = codec.Encode(full, cp)
        for c in range(cp.nChannels):
            self._chunk(cp, [(O[c], cp.nScaleBits)], A[c], S[c], M[c], H[c])
        return dict(joint=False, scaleFactor=S, bitAlloc=A, mantissa=M, overallScale=O, huffTable=H,
                    ms_switch=None, reservoir=cp.bitReservoir, tap=getattr(cp, "_tap", None))

    def JointWriteDataBlock(self, data, cp):
        """:793-972  channel 0 = M|L with the 4 overall scales and the ms bits, channel 1 = S|R."""
        full = [np.concatenate((cp.priorBlock[c], data[c])) for c in range(cp.nChannels)]
        cp.priorBlock = data
        cp.sfBands = sfbands_for(cp)
        S, A, M, O, ms, H = codec.JointEncode(full, cp)
        head0 = [(O[i], cp.nScaleBits) for i in range(4)] + [(ms[b], 1) for b in range(cp.sfBands.nBands)]
        self._chunk(cp, head0, A[0], S[0], M[0], H[0])
        self._chunk(cp, [], A[1], S[1], M[1], H[1])
        return dict(joint=True, scaleFactor=S, bitAlloc=A, mantissa=M, overallScale=O, huffTable=H,
                    ms_switch=ms, reservoir=cp.bitReservoir, tap=getattr(cp, "_tap", None))

    def Close(self, cp):
        """:973-984  flush: one extra NON-joint block of zeros (Q10)."""
        z = [np.zeros(cp.nMDCTLines, dtype=np.float64) for _ in range(cp.nChannels)]
        return self.WriteDataBlock(z, cp)

    def getvalue(self):
        return self.buf.getvalue()


class PACReader(object):
    def __init__(self, blob):
        self.fp = io.BytesIO(blob)
        self.params = self.ReadFileHeader()

    def ReadFileHeader(self):
        """:130-158"""
        if self.fp.read(4) != TAG:
            raise ValueError("Tried to read a non-PAC file into a PACFile object")
        sr, nCh, nSamp, nLinesM, nScale, nMant = unpack('<LHLLHH', self.fp.read(calcsize('<LHLLHH')))
        nBands = unpack('<L', self.fp.read(4))[0]
        nLines = unpack('<' + str(nBands) + 'H', self.fp.read(2 * nBands))
        cp = CodingParams()
        cp.sampleRate, cp.nChannels, cp.numSamples = sr, nCh, nSamp
        cp.nMDCTLines = cp.nSamplesPerBlock = nLinesM
        cp.nScaleBits, cp.nMantSizeBits = nScale, nMant
        cp.sfBands = ScaleFactorBands(nLines)
        cp.overlapAndAdd = [np.zeros(nLinesM, dtype=np.float64) for _ in range(nCh)]
        cp.a = cp.b = nLinesM
        cp.blkswBitA = cp.blkswBitB = 1
        return cp

    def _read_chunk_head(self, cp):
        s = self.fp.read(4)
        if not s:
            return None
        nBytes = unpack("<L", s)[0]
        pb = PackedBits()
        pb.SetPackedData(self.fp.read(nBytes))
        if pb.nBytes < nBytes:
            raise ValueError("Only read a partial block of coded PACFile data")
        huffTable = pb.ReadBits(N_HUFF_BITS)
        swA = pb.ReadBits(cp.blkswBitA)
        swB = pb.ReadBits(cp.blkswBitB)
        cp.a = (1 - swA) * cp.nMDCTLines + swA * 128
        cp.b = (1 - swB) * cp.nMDCTLines + swB * 128
        cp.sfBands = sfbands_for(cp)
        return pb, huffTable

    def _read_bands(self, pb, cp, huffTable):
        """band loop of :262-303 / :404-480: ba (stored ba-1), scale factor, then raw mantissas or a bit-serial
        prefix-code walk; the escape code is followed by ba raw bits.  Mantissas land at their line index."""
        sf = cp.sfBands
        bitAlloc, scaleFactor = [], []
        mant = np.zeros(cp.nMDCTLines, np.int32)
        T = None if huffTable == NO_TABLE else TABLES[huffTable]
        for b in range(sf.nBands):
            ba = pb.ReadBits(cp.nMantSizeBits)
            if ba:
                ba += 1
            bitAlloc.append(ba)
            scaleFactor.append(pb.ReadBits(cp.nScaleBits))
            if ba:
                lo = sf.lowerLine[b]
                for j in range(int(sf.nLines[b])):
                    if T is None:
                        mant[lo + j] = pb.ReadBits(ba)
                    else:
                        code = ""
                        while code not in T.rev:
                            code += "1" if pb.ReadBits(1) else "0"
                            if len(code) > T.maxlen:
                                raise ValueError("Something has gone horribly wrong...")
                        mant[lo + j] = pb.ReadBits(ba) if code == T.escape_code else T.rev[code]
        return scaleFactor, bitAlloc, mant

    def _ola(self, cp, decoded):
        out = []
        for c in range(cp.nChannels):
            out.append(np.add(cp.overlapAndAdd[c], decoded[c][:cp.a]))
            cp.overlapAndAdd[c] = decoded[c][cp.a:]
        return out

    def _eof(self, cp):
        if cp.overlapAndAdd:
            tail = cp.overlapAndAdd
            cp.overlapAndAdd = 0
            return tail
        return None

    def ReadDataBlock(self, cp):
        """:161-319"""
        decoded = []
        for c in range(cp.nChannels):
            h = self._read_chunk_head(cp)
            if h is None:
                return self._eof(cp)
            pb, huffTable = h
            overall = pb.ReadBits(cp.nScaleBits)
            S, A, M = self._read_bands(pb, cp, huffTable)
            decoded.append(codec.Decode(S, A, M, overall, cp))
        return self._ola(cp, decoded)

    def JointReadDataBlock(self, cp):
        """:321-585"""
        S, A, M, overall, ms = [], [], [], [], []
        for c in range(cp.nChannels):
            h = self._read_chunk_head(cp)
            if h is None:
                return self._eof(cp)
            pb, huffTable = h
            if c == 0:
                overall = [pb.ReadBits(cp.nScaleBits) for _ in range(4)]
                ms = [pb.ReadBits(1) for _ in range(cp.sfBands.nBands)]
            s, a, m = self._read_bands(pb, cp, huffTable)
            S.append(s); A.append(a); M.append(m)
        return self._ola(cp, codec.JointDecode(S, A, M, overall, cp, ms))
