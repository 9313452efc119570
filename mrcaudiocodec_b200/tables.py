"""Data-independent tables for libmrc.so, computed on the host with numpy exactly the way the reference computes
them on every call (so the values, including their last bits, are the reference's):

  kbd window      window.py:57-99   (KBDWindow, alpha=4; TransitionWindow(a=b=L) window.py:104-121 is the same vector)
  hann window     window.py:36-42
  MDCTFreq        psychoac.py:142-143
  Bark(f)         psychoac.py:29
  Thresh(f)       psychoac.py:23-25,  Intensity psychoac.py:18
  band table      psychoac.py:82-105 (25 Zwicker critical bands) / pacfileThem.py:216 (9-band short table)
  Huffman books   training_data/*_table.pkl, alphabetical order, shipped as huffman_tables.json
  block switching window.py:104-121 (TransitionWindow(a, b)), pacfileThem.py:213-219 (9 bands unless a+b = 2L),
                  :1146-1147 (transient detector's high-pass, scipy.signal like the reference), :1154 (thresholds)
"""
import json
import os

import numpy as np

HUFF_LUT = 65
_here = os.path.dirname(os.path.abspath(__file__))

CB_FREQ_LIMITS = [100, 200, 300, 400, 510, 630, 770, 920, 1080, 1270, 1480, 1720, 2000, 2320, 2700, 3150, 3700,
                  4400, 5300, 6400, 7700, 9500, 12000, 15500, 24000]
SHORT_FREQ_LIMITS = [300, 630, 1080, 1720, 2700, 4400, 7700, 15500, 24000]


def kbd_window(N, alpha=4.0):
    M = N / 2.0
    j = np.linspace(0, M, int(M + 1))
    j = np.square(np.divide(np.subtract(j, M / 2.0), M / 2.0))
    v = np.divide(np.i0(np.multiply(np.pi * alpha, np.sqrt(np.subtract(1.0, j)))), np.i0(np.pi * alpha))
    v2 = np.square(v)
    h = int(M)
    denom = np.sum(v2)
    # the reference sums through np.dot with triangular matrices (BLAS order); do the same
    top = np.sqrt(np.divide(np.dot(np.tril(np.ones((h, h))), v2[0:h]), denom))
    bot = np.sqrt(np.divide(np.dot(np.triu(np.ones((h, h))), v2[1:h + 1]), denom))
    return np.concatenate((top, bot))


def transition_window(a, b):
    """window.py:112-119: left half of KBD(2a) then right half of KBD(2b)."""
    return np.append(kbd_window(2 * a)[:a], kbd_window(2 * b)[b:])


def hann_window(N):
    n = np.add(np.linspace(0, N - 1, N), 0.5)
    return np.add(0.5, np.multiply(-0.5, np.cos(np.multiply((2.0 * np.pi) / N, n))))


def sine_window(N):
    n = np.add(np.linspace(0, N - 1, N), 0.5)
    return np.sin(np.multiply(np.pi / N, n))


def mdct_freqs(n_lines, sample_rate):
    return (np.arange(n_lines) + 0.5) * ((float(sample_rate) / n_lines) / 2.)


def bark(f):
    return 13 * np.arctan(0.76 * f / 1000.) + 3.5 * np.arctan((f / 7500.) ** 2)


def thresh(f):
    return (3.64 * ((f / 1000.) ** (-0.8))) - (6.5 * np.exp((-0.6 * (((f / 1000.) - 3.3) ** 2)))) \
        + ((10 ** (-3)) * ((f / 1000.) ** 4))


def intensity(spl):
    return 10 ** ((spl - 96) / 10)


def band_lines(n_lines, sample_rate, limits=None):
    limits = CB_FREQ_LIMITS if limits is None else limits
    f = mdct_freqs(n_lines, sample_rate)
    counts = np.zeros(len(limits), dtype=np.int64)
    i = j = 0
    while i < len(limits) - 1:
        while j < len(f) and f[j] < limits[i]:
            counts[i] += 1
            j += 1
        i += 1
    counts[i] = n_lines - counts.sum()
    return counts


def load_huffman():
    with open(os.path.join(_here, "huffman_tables.json")) as fh:
        d = json.load(fh)
    esc = np.zeros(len(d["tables"]), dtype=np.int32)
    lens = np.zeros((len(d["tables"]), HUFF_LUT), dtype=np.uint8)
    codes = np.zeros((len(d["tables"]), HUFF_LUT), dtype=np.uint16)
    for t, tab in enumerate(d["tables"]):
        esc[t] = tab["escape"]
        for v, c in tab["codes"].items():
            lens[t, int(v)] = len(c)
            codes[t, int(v)] = int(c, 2)
    return esc, lens, codes, [t["name"] for t in d["tables"]]


class Tables(object):
    """Everything mrc_set_tables needs, as contiguous numpy arrays kept alive by this object."""

    def __init__(self, n_mdct_lines, sample_rate, band_limits=None, window="kbd"):
        L = int(n_mdct_lines)
        self.band_nlines = np.ascontiguousarray(band_lines(L, sample_rate, band_limits), dtype=np.int32)
        # the analysis/synthesis window is a table to the kernels: KBD(alpha=4) is what the reference codec uses
        # (codecThem.py:316, :406-431 through TransitionWindow); the sine window (window.py:10-25) is the module's other
        # Princen-Bradley window, selectable here
        if window not in ("kbd", "sine"):
            raise ValueError("window must be 'kbd' or 'sine'")
        self.window_name = window
        self.kbd = np.ascontiguousarray(kbd_window(2 * L) if window == "kbd" else sine_window(2 * L), dtype=np.float64)
        self.hann = np.ascontiguousarray(hann_window(2 * L), dtype=np.float64)
        f = mdct_freqs(L, sample_rate)
        self.bark = np.ascontiguousarray(bark(f), dtype=np.float64)
        self.quiet = np.ascontiguousarray(intensity(thresh(f)), dtype=np.float64)
        self.huff_escape, self.huff_len, self.huff_code, self.huff_names = load_huffman()
        self.huff_len = np.ascontiguousarray(self.huff_len)
        self.huff_code = np.ascontiguousarray(self.huff_code)
        self.n_bands = int(len(self.band_nlines))
        self.band_lower = np.concatenate(([0], np.cumsum(self.band_nlines)[:-1])).astype(np.int64)


N_SHORT = 128                       # pacfileThem.py:1114
TRANSIENT_THRESHOLDS = (0.1, 0.075)  # pacfileThem.py:1154


def transient_sos(sample_rate):
    """pacfileThem.py:1146-1147, the reference's own scipy calls (scipy is the reference's dependency for exactly
    this).  tf2sos factors a 20th-order polynomial, so the sections' last bits depend on the LAPACK build; callers
    that need a bit-exact match with a stream encoded elsewhere pass that encoder's sections instead."""
    from scipy import signal
    b, a = signal.cheby2(20, 40, 9000. / sample_rate, 'high')
    return np.ascontiguousarray(signal.tf2sos(b, a), dtype=np.float64)


class BlockTables(object):
    """Tables of one block geometry of block switching: window halves a, b (each n_mdct_lines or 128)."""

    def __init__(self, a, b, n_mdct_lines, sample_rate):
        self.a, self.b = int(a), int(b)
        half = (self.a + self.b) // 2
        limits = None if self.a + self.b == 2 * n_mdct_lines else SHORT_FREQ_LIMITS      # pacfileThem.py:208-219
        self.band_nlines = np.ascontiguousarray(band_lines(half, sample_rate, limits), dtype=np.int32)
        self.window = np.ascontiguousarray(transition_window(self.a, self.b), dtype=np.float64)
        self.hann = np.ascontiguousarray(hann_window(self.a + self.b), dtype=np.float64)
        f = mdct_freqs(half, sample_rate)
        self.bark = np.ascontiguousarray(bark(f), dtype=np.float64)
        self.quiet = np.ascontiguousarray(intensity(thresh(f)), dtype=np.float64)
        self.n_bands = int(len(self.band_nlines))
        self.n_lines = half
        self.band_lower = np.concatenate(([0], np.cumsum(self.band_nlines)[:-1])).astype(np.int64)
