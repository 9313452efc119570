"""The reference's own known-answer tests, run against the oracle restatement."""
import math

import numpy as np

import mrc_oracle as o


def test_tdac_12_sample_vector():
    """mdct.py:131,174-179 -- 0.5*IMDCT(MDCT(.)) overlap-added reproduces x after a half-block delay."""
    x = np.array([3, 3, 3, 3, 2, 0, -2, -4, -1, 0, 1, 2.])
    z = np.zeros(4)
    prev = np.zeros(8)
    out = []
    for k in range(4):
        cur = x[k * 4:(k + 1) * 4] if k < 3 else z
        pr = x[(k - 1) * 4:k * 4] if k > 0 else z
        blk = np.concatenate([pr, cur])
        im = 0.5 * o.mdct.IMDCT(o.mdct.MDCT(blk, 4, 4), 4, 4)
        im_slow = 0.5 * o.mdct.MDCTslow(o.mdct.MDCTslow(blk, 4, 4), 4, 4, True)
        np.testing.assert_array_almost_equal(im, im_slow)
        out += list(im[:4] + prev[4:])
        prev = im
    assert np.rint(out).astype(int).tolist() == [0, 0, 0, 0, 3, 3, 3, 3, 2, 0, -2, -4, -1, 0, 1, 2]


def test_fast_vs_slow_mdct_1024():
    """mdct.py:184-210"""
    x = np.arange(1024.)
    X = o.mdct.MDCT(x, 512, 512)
    np.testing.assert_array_almost_equal(X, o.mdct.MDCTslow(x, 512, 512))
    np.testing.assert_array_almost_equal(o.mdct.IMDCT(X, 512, 512), o.mdct.MDCTslow(X, 512, 512, True))


def test_bitpack_vector():
    """bitpack.py:183-196 -- (3,5,11,3,1) in (4,3,5,3,1) bits -> 0x3A 0xB7, and back."""
    vals, lens = (3, 5, 11, 3, 1), (4, 3, 5, 3, 1)
    bp = o.bitpack.PackedBits()
    bp.Size(2)
    for v, n in zip(vals, lens):
        bp.WriteBits(v, n)
    assert bp.GetPackedData() == bytes([0x3A, 0xB7])
    bp2 = o.bitpack.PackedBits()
    bp2.SetPackedData(bp.GetPackedData())
    assert [bp2.ReadBits(n) for n in lens] == list(vals)


def test_log2_floor_is_exact_here():
    """quantize.py:131 uses int(math.log(code, 2)); the kernels use 31-clz.  Equal for every reachable code
    on this libm (SURVEY.md §7 'hard parts')."""
    for k in range(0, 31):
        for c in (2 ** k - 1, 2 ** k, 2 ** k + 1):
            if c >= 1 and c < 2 ** 30:
                assert int(math.log(c, 2)) == c.bit_length() - 1, c
    rng = np.random.default_rng(0)
    for c in rng.integers(1, 2 ** 30, 20000):
        assert int(math.log(int(c), 2)) == int(c).bit_length() - 1


def test_band_tables():
    """psychoac.py:86-105 at 48 kHz / 1024 lines and the 9-band short table."""
    n = o.psychoac.AssignMDCTLinesFromFreqLimits(1024, 48000).astype(int).tolist()
    assert n == [4, 5, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 12, 14, 16, 19, 24, 30, 38, 47, 56, 76, 107, 149, 363]
    s = o.psychoac.AssignMDCTLinesFromFreqLimits(128, 48000, o.psychoac.shortFreqLimits).astype(int).tolist()
    assert s == [2, 1, 3, 3, 5, 9, 18, 42, 45]


def test_pcm_conversion_quirks():
    """pcmfile.py:93-99 + quantize.py:95-107: -32768 -> 0.0 (Q1); round trip of every other code is exact."""
    c = np.arange(-32768, 32768).astype(np.int16)
    x = o.pcm.pcm_to_fraction(c)
    assert x[0] == 0.0
    assert x[32768 + 32767] == 2 * 32767 / 65535.0
    back = o.pcm.fraction_to_pcm(x)
    assert np.array_equal(back[1:], c[1:])
    assert o.pcm.fraction_to_pcm(np.array([1.5, -1.5, 1.0, -1.0])).tolist() == [32767, -32767, 32767, -32767]


def test_huffman_tables_are_complete_prefix_codes():
    for T in o.tables.TABLES:
        assert abs(sum(2.0 ** -len(c) for c in T.codes.values()) - 1.0) < 1e-12
        assert T.escape in T.codes
    assert [T.name for T in o.tables.TABLES] == ["percussive", "silence", "speech", "tonal"]
