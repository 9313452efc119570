"""
ref_driver.py -- drive the shimmed, otherwise UNMODIFIED reference (oracle/ref_shim.py) through the canonical
file flow, in memory.  TEST INFRASTRUCTURE ONLY; works only where /root/reference exists.

encode: WAV bytes -> reference PCMFile.ReadFileHeader/ReadDataBlock (pcmfile.py:34-102) ->
        PACFile.WriteFileHeader + JointWriteDataBlock|WriteDataBlock per block + Close (pacfileThem.py:586-984)
decode: PACFile.ReadFileHeader + JointReadDataBlock|ReadDataBlock (+ ReadDataBlock for the flush pair, Q10) ->
        PCMFile.WriteDataBlock (pcmfile.py:156-185)
The reference's own seam methods PACFile.Encode/JointEncode are wrapped to record what they return.
"""
import io
import struct

import numpy as np

import ref_shim


class _MemFile(io.BytesIO):
    def __init__(self, data=b"", mode="rb"):
        io.BytesIO.__init__(self, data)
        self.mode = mode

    def close(self):          # keep the buffer readable after the reference closes it
        pass


def wav_bytes(pcm, sampleRate):
    pcm = np.asarray(pcm, dtype='<i2')
    n, nCh = pcm.shape
    data = pcm.tobytes()
    return struct.pack('<4sL4s4sLHHLLHH4sL', b"RIFF", 36 + len(data), b"WAVE", b"fmt ", 16, 1, nCh, sampleRate,
                       sampleRate * nCh * 2, nCh * 2, 16, b"data", len(data)) + data


def ref_encode(pcm, sampleRate=48000, joint=True, nMDCTLines=1024, nScaleBits=4, nMantSizeBits=4,
               targetBitsPerSample=128000. / 48000., trace=True):
    m = ref_shim.load()
    PCMFile, PACFile = m["pcmfile"].PCMFile, m["pacfileThem"].PACFile
    with ref_shim.in_reference_cwd():
        inF = PCMFile("<mem>.wav")
        inF.fp = _MemFile(wav_bytes(pcm, sampleRate))
        cp = inF.ReadFileHeader()
        cp.nMDCTLines = nMDCTLines
        cp.nScaleBits = nScaleBits
        cp.nMantSizeBits = nMantSizeBits
        cp.targetBitsPerSample = targetBitsPerSample
        cp.nSamplesPerBlock = cp.nMDCTLines
        cp.bitReservoir = 0
        cp.nSamplesShort = 128
        cp.a = cp.b = cp.nMDCTLines
        cp.blkswBitA = cp.blkswBitB = 1
        outF = PACFile("<mem>.pac")
        outF.fp = _MemFile(mode="wb")
        outF.WriteFileHeader(cp)
        blocks = []
        rec = {}

        def wrap(name):
            orig = getattr(outF, name)

            def f(*a, **k):
                r = orig(*a, **k)
                rec["ret"] = r
                return r
            setattr(outF, name, f)
        wrap("Encode")
        wrap("JointEncode")

        def snap(is_joint):
            r = rec["ret"]
            if is_joint:
                S, A, M, O, ms, H = r
            else:
                S, A, M, O, H = r
                ms = None
            blocks.append(dict(joint=is_joint, scaleFactor=[np.array(s) for s in S],
                               bitAlloc=[np.array(a) for a in A],
                               mantissa=[list(x) if not isinstance(x, np.ndarray) else x.copy() for x in M],
                               overallScale=list(O), huffTable=list(H),
                               ms_switch=None if ms is None else list(ms), reservoir=int(cp.bitReservoir)))
        blocky = 0
        while True:
            data = inF.ReadDataBlock(cp, blocky)
            if not data:
                break
            if joint:
                outF.JointWriteDataBlock(data, cp, blocky)
            else:
                outF.WriteDataBlock(data, cp, blocky)
            if trace:
                snap(joint)
            blocky += 1
        outF.Close(cp)
        if trace:
            snap(False)
        return outF.fp.getvalue(), blocks


def ref_encode_switched(pcm, sampleRate=48000, nMDCTLines=1024, nScaleBits=4, nMantSizeBits=4,
                        targetBitsPerSample=128000. / 48000., drop_last=False):
    """The reference's `__main__` encode loop with block switching (pacfileThem.py:1142-1226), restated around the
    reference's OWN TransientDetector / JointWriteDataBlock / Close.  drop_last=True is the shipped behaviour
    (the last block read is never written, Q11); drop_last=False is the canonical driver: the last block is
    written with no look-ahead information and Close() runs with b = nMDCTLines.
    Returns (pac bytes, [(a, b) per written block], [detections per PCM block])."""
    from scipy import signal
    m = ref_shim.load()
    PCMFile, PACFile = m["pcmfile"].PCMFile, m["pacfileThem"].PACFile
    TransientDetector = m["pacfileThem"].TransientDetector
    with ref_shim.in_reference_cwd():
        inF = PCMFile("<mem>.wav")
        inF.fp = _MemFile(wav_bytes(pcm, sampleRate))
        cp = inF.ReadFileHeader()
        cp.nMDCTLines = nMDCTLines
        cp.nScaleBits = nScaleBits
        cp.nMantSizeBits = nMantSizeBits
        cp.targetBitsPerSample = targetBitsPerSample
        cp.nSamplesPerBlock = cp.nMDCTLines
        cp.bitReservoir = 0
        cp.nSamplesShort = 128
        cp.a = cp.b = cp.nMDCTLines
        cp.blkswBitA = cp.blkswBitB = 1
        outF = PACFile("<mem>.pac")
        outF.fp = _MemFile(mode="wb")
        outF.WriteFileHeader(cp)
        b_, a_ = signal.cheby2(20, 40, 9000. / cp.sampleRate, 'high')          # :1146-1147
        sos = signal.tf2sos(b_, a_)
        nSeg = cp.nSamplesPerBlock // cp.nSamplesShort
        cp.P = np.zeros((cp.nChannels, 1 + nSeg))
        T = np.array([0.1, 0.075])
        geom, dets = [], []
        blocky = 0
        firstBlock = True
        dataMem = blkswMem = None

        def write(dataMem, blkswMem, blksw):
            nonlocal blocky
            if np.sum(blkswMem) > 1 or (blksw is not None and any(blksw == 1)):       # :1192
                for i in range(nSeg):
                    cp.b = cp.nSamplesShort
                    outF.JointWriteDataBlock(dataMem[:, cp.b * i:cp.b * (i + 1)], cp, blocky)
                    geom.append((cp.a, cp.b))
                    cp.a = cp.b
                    blocky += 1
            else:
                cp.b = cp.nSamplesPerBlock
                outF.JointWriteDataBlock(dataMem, cp, blocky)
                geom.append((cp.a, cp.b))
                cp.a = cp.b
                blocky += 1

        while True:
            data = inF.ReadDataBlock(cp, blocky)
            if not data:
                break
            data = np.vstack(data)
            blksw = TransientDetector(data, cp, sos, T)
            dets.append(blksw)
            if firstBlock:
                dataMem, blkswMem, firstBlock = data, blksw, False
                continue
            write(dataMem, blkswMem, blksw)
            dataMem, blkswMem = data, blksw
        if not drop_last and dataMem is not None:
            write(dataMem, blkswMem, None)
            cp.b = cp.nMDCTLines
        outF.Close(cp)
        geom.append((cp.a, cp.b))
        return outF.fp.getvalue(), geom, dets


def ref_decode(blob, joint=True):
    m = ref_shim.load()
    PCMFile, PACFile = m["pcmfile"].PCMFile, m["pacfileThem"].PACFile
    import mrc_oracle.driver as drv
    with ref_shim.in_reference_cwd():
        inF = PACFile("<mem>.pac")
        inF.fp = _MemFile(blob)
        cp = inF.ReadFileHeader()
        cp.bitsPerSample = 16
        cp.nSamplesShort = 128
        cp.a = cp.b = cp.nMDCTLines
        cp.blkswBitA = cp.blkswBitB = 1
        nPairs = drv.count_block_pairs(blob, cp.nChannels)
        outF = PCMFile("<mem>.wav")
        outF.fp = _MemFile(mode="wb")
        first = True
        i = 0
        while True:
            if joint and i != nPairs - 1:
                data = inF.JointReadDataBlock(cp, i)
            else:
                data = inF.ReadDataBlock(cp, i)
            i += 1
            if not data:
                break
            data = np.vstack(data)
            if first:
                first = False
                continue
            outF.WriteDataBlock([data[c] for c in range(data.shape[0])], cp, i)
        raw = outF.fp.getvalue()
        return np.frombuffer(raw, dtype='<i2').reshape(-1, cp.nChannels).copy()
