"""Psychoacoustic model.  Follows /root/reference/psychoac.py: SPL :8-12, Intensity :14-18, Thresh :20-25,
Bark :27-29, Masker :31-78, cbFreqLimits :82-84, AssignMDCTLinesFromFreqLimits :86-105,
ScaleFactorBands :107-131, getMaskedThreshold :134-173, CalcSMRs :176-219."""
import numpy as np
from .window import HanningWindow


def SPL(intensity):
    """:12  max(96 + 10*log10(I), -30).  log10(0) = -inf is expected for silent input."""
    with np.errstate(divide="ignore"):
        return np.maximum(96 + 10 * np.log10(intensity), -30)


def Intensity(spl):
    """:18"""
    return 10 ** ((spl - 96) / 10)


def Thresh(f):
    """:23-25  threshold in quiet, dB SPL."""
    return (3.64 * ((f / 1000.) ** (-0.8))) \
        - (6.5 * np.exp((-0.6 * (((f / 1000.) - 3.3) ** 2)))) \
        + ((10 ** (-3)) * ((f / 1000.) ** 4))


def Bark(f):
    """:29"""
    return 13 * np.arctan(0.76 * f / 1000.) + 3.5 * np.arctan((f / 7500.) ** 2)


class Masker(object):
    """:31-78  tonal masker (drop 14.5+0.5 = 15 dB); noise maskers are never instantiated by the codec."""

    def __init__(self, f, spl, isTonal=True):
        self.drop = 14.5 + 0.5 if isTonal else 5.5
        self.z = Bark(f)
        self.SPL = spl
        self.f = f

    def vIntensityAtBark(self, zVec):
        """:68-78  evaluation order of the reference expression is kept."""
        dz = zVec - self.z
        sign = dz > 0.5
        mag = np.abs(dz) > 0.5
        return Intensity(self.SPL - self.drop
                         + -27 * (np.abs(dz) - 0.5) * mag
                         + 0.37 * np.maximum(self.SPL - 40, 0) * (np.abs(dz) - 0.5) * mag * sign)


cbFreqLimits = [100, 200, 300, 400, 510, 630, 770, 920, 1080,
                1270, 1480, 1720, 2000, 2320, 2700, 3150, 3700,
                4400, 5300, 6400, 7700, 9500, 12000, 15500, 24000]
shortFreqLimits = [300, 630, 1080, 1720, 2700, 4400, 7700, 15500, 24000]     # pacfileThem.py:216


def AssignMDCTLinesFromFreqLimits(nMDCTLines, sampleRate, flimit=cbFreqLimits):
    """:86-105  count lines whose centre frequency is below each limit; the last band takes the remainder."""
    f = (np.arange(nMDCTLines) + 0.5) * ((float(sampleRate) / nMDCTLines) / 2.)
    counts = np.zeros(len(flimit))
    i = j = 0
    while i < len(flimit) - 1:
        while j < len(f) and f[j] < flimit[i]:
            counts[i] += 1
            j += 1
        i += 1
    counts[i] = nMDCTLines - sum(counts)
    return counts


class ScaleFactorBands(object):
    """:107-131"""

    def __init__(self, nLines):
        self.nBands = len(nLines)
        self.lowerLine = np.cumsum(np.append([0], nLines[0:self.nBands - 1]), dtype=int)
        self.upperLine = np.cumsum(np.transpose(nLines), dtype=int) - 1
        self.nLines = self.upperLine - self.lowerLine + 1


def find_peaks(XI, N):
    """:156-171  strict local maxima of XI at bins 1 .. N/2-102 (loop index i runs 2 .. N/2-101)."""
    hi = N // 2 - 100
    c = XI[1:hi - 1]
    return np.nonzero((c > XI[0:hi - 2]) & (c > XI[2:hi]))[0] + 1


def getMaskedThreshold(data, MDCTdata, MDCTscale, sampleRate, sfBands, return_peaks=False):
    """:134-173.  Hann window, FFT, intensity, tonal peaks, spreading summed onto the threshold in quiet in
    ascending peak order.  (sampleRate//N) is the reference's Python-2 integer division, quirk Q2."""
    nLines = len(MDCTdata)
    MDCTFreq = (np.arange(nLines) + 0.5) * ((float(sampleRate) / nLines) / 2.)
    N = len(data)
    X = np.fft.fft(HanningWindow(data))
    XI = 4. * (np.abs(X) ** 2.) / ((N ** 2.) * (3. / 8.))
    totalMask = Intensity(Thresh(MDCTFreq))
    zVec = Bark(MDCTFreq)
    peaks = find_peaks(XI, N)
    plist = []
    for p in peaks:
        XI0, XI1, XI2 = XI[p - 1], XI[p], XI[p + 1]
        XMask = SPL(XI0 + XI1 + XI2)
        nMask = (int(sampleRate) // N) * ((p - 1) * XI0 + p * XI1 + (p + 1) * XI2) / (XI0 + XI1 + XI2)
        tone = Masker(nMask, XMask)
        totalMask += tone.vIntensityAtBark(zVec)
        if return_peaks:
            plist.append((int(p), float(XMask), float(tone.z)))
    thr = SPL(totalMask)
    return (thr, plist) if return_peaks else thr


def CalcSMRs(data, MDCTdata, MDCTscale, sampleRate, sfBands, ms=0, preCalcThresh=0.0):
    """:176-219.  The reference ignores ms/preCalcThresh (the threshold is unconditionally recomputed at :210,
    quirk Q7) and evaluates getMaskedThreshold twice with identical inputs (Q8); once is bit-identical."""
    maskThresh = getMaskedThreshold(data, MDCTdata, MDCTscale, sampleRate, sfBands)
    MDCTSPL = SPL(2. * (np.abs(MDCTdata) ** 2.) / (1. / 2.)) - 6. * MDCTscale
    d = MDCTSPL - maskThresh
    SMR = np.zeros(sfBands.nBands)
    for i in range(sfBands.nBands):
        SMR[i] = np.amax(d[sfBands.lowerLine[i]:sfBands.upperLine[i] + 1])
    return SMR
