"""M/S stereo helpers.  Follows /root/reference/ms_stereo.py (MSSwitchSFBands :5-27, ReconstructLR :33-49,
OverallSMRs :70-81).  StereoMaskingFactor (:53-67) is dead code in the reference (its result never reaches
the output) and is not restated."""
import numpy as np


def MSSwitchSFBands(mdct_left, mdct_right, sfBands):
    """:5-27  ms=1 iff sum|l^2-r^2| < 0.8*sum|l^2+r^2| over the band (unscaled MDCT lines)."""
    d = np.square(mdct_left) - np.square(mdct_right)
    s = np.square(mdct_left) + np.square(mdct_right)
    out = []
    for i in range(sfBands.nBands):
        lo, hi = sfBands.lowerLine[i], sfBands.upperLine[i] + 1
        out.append(1 if np.sum(np.abs(d[lo:hi])) < 0.8 * np.sum(np.abs(s[lo:hi])) else 0)
    return out


def ReconstructLR(m1, m2, sfBands, ms_switch):
    """:33-49  L=M+S, R=M-S on ms bands, pass-through elsewhere."""
    left = np.array(m1, dtype=np.float64, copy=True)
    right = np.array(m2, dtype=np.float64, copy=True)
    for i in range(sfBands.nBands):
        if ms_switch[i] == 1:
            lo, hi = sfBands.lowerLine[i], sfBands.upperLine[i] + 1
            left[lo:hi] = m1[lo:hi] + m2[lo:hi]
            right[lo:hi] = m1[lo:hi] - m2[lo:hi]
    return left, right


def OverallSMRs(SMR_l, SMR_r, SMR_m, SMR_s, sfBands, ms_switch):
    """:70-81  per band pick (M,S) if ms else (L,R)."""
    a = [SMR_m[i] if ms_switch[i] == 1 else SMR_l[i] for i in range(sfBands.nBands)]
    b = [SMR_s[i] if ms_switch[i] == 1 else SMR_r[i] for i in range(sfBands.nBands)]
    return a, b
