"""Greedy water-filling bit allocation.  Follows /root/reference/bitalloc.py:106-155 (BitAlloc, the only
allocator the codec calls).  Quirk Q5 kept: the first grant gives 2 bits for 2*nLines but only checks
nLines <= bitsLeft, so bitsLeft can go negative."""
import numpy as np


def BitAlloc(bitBudget, maxMantBits, nBands, nLines, SMR):
    smr = np.array(SMR, dtype=np.float64)      # the reference works on (and clobbers) the caller's array
    bitsLeft = bitBudget                       # python float
    excluded = 0
    bits = np.zeros(nBands)
    while bitsLeft > 0:
        i = int(np.argmax(smr))                # first maximum wins
        if bits[i] < maxMantBits and nLines[i] <= bitsLeft:
            if bits[i] == 0:
                bits[i] += 2
                bitsLeft -= 2 * nLines[i]
                smr[i] -= 12.0
            else:
                bits[i] += 1
                bitsLeft -= nLines[i]
                smr[i] -= 6.0
        else:
            smr[i] = -99999999999999999.0
            excluded += 1
            if excluded == nBands:
                break
    return bits, int(bitsLeft)                 # int() truncates toward zero
