"""Host-side helpers for the .pac container as the shipped reference writes it (pacfileThem.py:586-613 header,
<L nBytes-prefixed channel chunks).  Pure byte bookkeeping -- no codec arithmetic."""
import struct

import numpy as np


def parse_header(blob):
    if blob[:4] != b'PAC ':
        raise ValueError("not a PAC file")
    sr, nch, nsamp, L, nsb, nmsb = struct.unpack('<LHLLHH', blob[4:22])
    nb = struct.unpack('<L', blob[22:26])[0]
    nlines = struct.unpack('<%dH' % nb, blob[26:26 + 2 * nb])
    return dict(sampleRate=sr, nChannels=nch, numSamples=nsamp, nMDCTLines=L, nScaleBits=nsb, nMantSizeBits=nmsb,
                nBands=nb, nLines=list(nlines), headerBytes=26 + 2 * nb)


def chunk_index(blob):
    """[(payload offset, nBytes)] of every channel chunk."""
    h = parse_header(blob)
    pos = h["headerBytes"]
    out = []
    while pos < len(blob):
        n = struct.unpack('<L', blob[pos:pos + 4])[0]
        out.append((pos + 4, n))
        pos += 4 + n
    return out


def huff_table_ids(blob):
    """table id (first 4 bits) of every channel chunk."""
    return np.array([blob[o] >> 4 for o, n in chunk_index(blob)], dtype=np.int32)
