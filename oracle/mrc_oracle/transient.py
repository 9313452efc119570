"""Block switching: transient detector and the look-ahead decision.  Follows /root/reference/pacfileThem.py:
TransientDetector :1021-1056, the filter design :1146-1147 (20th-order Chebyshev-II high-pass through
scipy.signal.cheby2 + tf2sos, exactly the reference's calls), thresholds :1154, and the decision of the
`__main__` loop :1182-1214.

The reference's loop has a one-block look-ahead: the block read one iteration earlier (`dataMem`) is written as
eight 128-sample short blocks iff  sum(blkswMem) > 1  or  any(blksw == 1)  (:1192), i.e. iff its own detection
holds a transient position >= 2 (positions are unique integers 1..8, so their sum exceeds 1 exactly then) or
the NEXT block has one in its first 128 samples.  The reference never writes the last block it read (Q11);
the canonical driver (driver.py) writes it too, with no look-ahead information."""
import numpy as np
from scipy import signal

N_SHORT = 128
THRESHOLDS = np.array([0.1, 0.075])          # pacfileThem.py:1154


def design_sos(sampleRate):
    """pacfileThem.py:1146-1147"""
    b, a = signal.cheby2(20, 40, 9000. / sampleRate, 'high')
    return signal.tf2sos(b, a)


def sosfilt_plain(sos, x):
    """What scipy.signal.sosfilt computes (direct form II transposed, products and sums unfused, in this
    order) -- bit-identical to the library call; the CUDA kernel follows this loop."""
    ns = sos.shape[0]
    z = np.zeros((ns, 2))
    out = np.empty(len(x), dtype=np.float64)
    c = [[float(v) for v in row] for row in sos]
    for n in range(len(x)):
        xc = float(x[n])
        for s in range(ns):
            b0, b1, b2, _, a1, a2 = c[s]
            xn = b0 * xc + z[s, 0]
            z[s, 0] = (b1 * xc - a1 * xn) + z[s, 1]
            z[s, 1] = b2 * xc - a2 * xn
            xc = xn
        out[n] = xc
    return out


def TransientDetector(data, codingParams, sos, T):
    """:1021-1056.  data [nChannels][nSamplesPerBlock]; codingParams.P [nChannels][1 + nSegments] carries the
    previous block's last segment peak in column 0.  Returns the sorted unique transient positions (1-based
    128-sample segments) as a float array, like the reference."""
    cp = codingParams
    nSeg = cp.nSamplesPerBlock // cp.nSamplesShort
    blksw = np.array([])
    for iCh in range(cp.nChannels):
        dataFilt = signal.sosfilt(sos, data[iCh])          # zero initial state on every call
        for i in range(nSeg):
            cp.P[iCh][i + 1] = np.amax(np.abs(dataFilt[i * cp.nSamplesShort:(i + 1) * cp.nSamplesShort]))
        if np.amax(np.abs(dataFilt)) > T[0]:
            for i in range(nSeg):
                if cp.P[iCh][i + 1] * T[1] > cp.P[iCh][i]:
                    blksw = np.append(blksw, i + 1)
    cp.P[:, 0] = cp.P[:, nSeg]
    return np.unique(blksw[np.nonzero(blksw)])


def wants_short(blkswMem, blksw):
    """:1192  (blksw = detection of the following block, or None when there is none)"""
    return bool(np.sum(blkswMem) > 1 or (blksw is not None and np.any(blksw == 1)))
