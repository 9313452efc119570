"""Canonical file-level flow: header, one (Joint)WriteDataBlock per
nMDCTLines-frame PCM block (the last one zero padded, pcmfile.py:79-82), Close() = one extra NON-joint block
of zeros.  Decode: (Joint)ReadDataBlock for every block pair but the last, ReadDataBlock for the flush pair,
first decoded block dropped (pacfileThem.py:1175-1177), saved overlap returned once at EOF (:178-185).
encode_pcm is the plain per-block loop of audiofile.py:24-38 (long blocks only); encode_pcm_switched is the
reference's `__main__` loop with its transient detector and one-block look-ahead (pacfileThem.py:1142-1215,
SURVEY.md §8 f1), made canonical in two places where the shipped loop is broken (Q11): the last block read is
written too (with no look-ahead information), and the Close() flush block always has b = nMDCTLines."""
from struct import unpack

import numpy as np

from .pacfile import CodingParams, PACWriter, PACReader
from .pcm import pcm_to_fraction, fraction_to_pcm
from . import transient


def make_params(sampleRate=48000, nChannels=2, numSamples=0, nMDCTLines=1024, nScaleBits=4, nMantSizeBits=4,
                targetBitsPerSample=128000. / 48000.):
    """pacfileThem.py:1105-1121 (the reference hard-codes 2.86 bits/sample; BASELINE uses 128 kb/s/ch)."""
    cp = CodingParams()
    cp.sampleRate = int(sampleRate)
    cp.nChannels = nChannels
    cp.numSamples = int(numSamples)
    cp.bitsPerSample = 16
    cp.nMDCTLines = cp.nSamplesPerBlock = nMDCTLines
    cp.nScaleBits = nScaleBits
    cp.nMantSizeBits = nMantSizeBits
    cp.targetBitsPerSample = targetBitsPerSample
    cp.bitReservoir = 0
    cp.nSamplesShort = 128
    cp.a = cp.b = nMDCTLines
    cp.blkswBitA = cp.blkswBitB = 1
    return cp


def encode_pcm(pcm, joint=True, trace=False, **kw):
    """pcm: int16 array [numSamples, nChannels] (interleaved frames).  Returns (pac bytes, per-block trace)."""
    pcm = np.asarray(pcm, dtype=np.int16)
    n, nCh = pcm.shape
    cp = make_params(numSamples=n, nChannels=nCh, **kw)
    w = PACWriter(cp)
    L = cp.nMDCTLines
    nBlocks = (n + L - 1) // L
    blocks = []
    for b in range(nBlocks):
        seg = pcm[b * L:(b + 1) * L]
        if seg.shape[0] < L:
            seg = np.concatenate((seg, np.zeros((L - seg.shape[0], nCh), dtype=np.int16)))
        data = [pcm_to_fraction(seg[:, c]) for c in range(nCh)]
        r = w.JointWriteDataBlock(data, cp) if joint else w.WriteDataBlock(data, cp)
        if trace:
            blocks.append(r)
    r = w.Close(cp)
    if trace:
        blocks.append(r)
    return w.getvalue(), blocks


def encode_window(pcm_prior, pcm_blocks, reservoir_in, joint=True, **kw):
    """Blocks k .. k+n-1 out of the middle of a long-block stream: what (Joint)WriteDataBlock (pacfileThem.py:622-790,
    :793-972) writes for them when the block before left `pcm_prior` in codingParams.priorBlock (:631, :802) and
    `reservoir_in` in codingParams.bitReservoir (codecThem.py:308, :391) -- the only two things a block inherits.
    pcm_prior: int16 [nMDCTLines, nCh]; pcm_blocks: int16 [n*nMDCTLines, nCh].
    Returns (list of n byte strings: both channel chunks of a block with their <L prefixes, list of n reservoirs)."""
    pcm_blocks = np.asarray(pcm_blocks, dtype=np.int16)
    nCh = pcm_blocks.shape[1]
    cp = make_params(numSamples=pcm_blocks.shape[0], nChannels=nCh, **kw)
    w = PACWriter(cp)
    L = cp.nMDCTLines
    assert pcm_blocks.shape[0] % L == 0 and np.asarray(pcm_prior).shape == (L, nCh)
    cp.priorBlock = [pcm_to_fraction(np.asarray(pcm_prior, dtype=np.int16)[:, c]) for c in range(nCh)]
    cp.bitReservoir = int(reservoir_in)
    chunks, res = [], []
    for b in range(pcm_blocks.shape[0] // L):
        pos = w.buf.tell()
        data = [pcm_to_fraction(pcm_blocks[b * L:(b + 1) * L, c]) for c in range(nCh)]
        if joint:
            w.JointWriteDataBlock(data, cp)
        else:
            w.WriteDataBlock(data, cp)
        chunks.append(w.buf.getvalue()[pos:])
        res.append(int(cp.bitReservoir))
    return chunks, res


def encode_pcm_switched(pcm, trace=False, sos=None, **kw):
    """pacfileThem.py:1142-1215 with block switching: every nMDCTLines-frame block goes through TransientDetector;
    block k is written as 8 short blocks (b = 128 each) iff wants_short(detection of k, detection of k+1), else as
    one long block; a = the previous written block's b.  Always the joint flow, like the reference's loop.
    Returns (pac bytes, per-written-block trace, list of (a, b) per written block incl. the flush block)."""
    pcm = np.asarray(pcm, dtype=np.int16)
    n, nCh = pcm.shape
    cp = make_params(numSamples=n, nChannels=nCh, **kw)
    w = PACWriter(cp)
    L = cp.nMDCTLines
    nSeg = L // cp.nSamplesShort
    if sos is None:
        sos = transient.design_sos(cp.sampleRate)
    cp.P = np.zeros((nCh, 1 + nSeg))
    nBlocks = (n + L - 1) // L
    data, det = [], []
    for b in range(nBlocks):
        seg = pcm[b * L:(b + 1) * L]
        if seg.shape[0] < L:
            seg = np.concatenate((seg, np.zeros((L - seg.shape[0], nCh), dtype=np.int16)))
        d = np.vstack([pcm_to_fraction(seg[:, c]) for c in range(nCh)])
        data.append(d)
        det.append(transient.TransientDetector(d, cp, sos, transient.THRESHOLDS))
    blocks, geom = [], []
    for b in range(nBlocks):
        if transient.wants_short(det[b], det[b + 1] if b + 1 < nBlocks else None):
            for i in range(nSeg):
                cp.b = cp.nSamplesShort
                r = w.JointWriteDataBlock([data[b][c][cp.b * i:cp.b * (i + 1)] for c in range(nCh)], cp)
                geom.append((cp.a, cp.b))
                cp.a = cp.b
                if trace:
                    blocks.append(r)
        else:
            cp.b = L
            r = w.JointWriteDataBlock([data[b][c] for c in range(nCh)], cp)
            geom.append((cp.a, cp.b))
            cp.a = cp.b
            if trace:
                blocks.append(r)
    cp.b = L
    r = w.Close(cp)
    geom.append((cp.a, cp.b))
    if trace:
        blocks.append(r)
    return w.getvalue(), blocks, geom, det


def count_block_pairs(blob, nChannels):
    """walk the <L nBytes> chain after the header."""
    nBands = unpack('<L', blob[22:26])[0]
    pos = 26 + 2 * nBands
    n = 0
    while pos < len(blob):
        pos += 4 + unpack('<L', blob[pos:pos + 4])[0]
        n += 1
    return n // nChannels


def decode_pac(blob, joint=True):
    """Returns int16 PCM [nBlocks*nMDCTLines, nChannels]: pairs 1..B overlap-added with their predecessor plus
    the final saved tail (B+1 pairs in the file -> B+1 PCM blocks; the last one is the flush tail)."""
    r = PACReader(blob)
    cp = r.params
    cp.bitsPerSample = 16
    nPairs = count_block_pairs(blob, cp.nChannels)
    out = []
    first = True
    i = 0
    while True:
        if joint and i != nPairs - 1:
            data = r.JointReadDataBlock(cp)
        else:
            data = r.ReadDataBlock(cp)
        i += 1
        if not data:
            break
        if first:
            first = False
            continue
        out.append(np.stack([fraction_to_pcm(np.array(d)) for d in data], axis=1))
    return np.concatenate(out, axis=0) if out else np.zeros((0, cp.nChannels), np.int16)
