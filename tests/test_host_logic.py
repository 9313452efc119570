"""CPU-side checks of the product package: the C ABI loads and exports every symbol include/mrc.h declares, the
host tables equal the oracle's, the product path fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mrc.h")).read()
    declared = set(re.findall(r'\b(mrc_[a-z0-9_]+)\s*\(', hdr))
    declared -= {"mrc_ctx", "mrc_config", "mrc_tables"}
    so = os.path.join(ROOT, "mrcaudiocodec_b200", "libmrc.so")
    if not os.path.exists(so):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(so)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    from mrcaudiocodec_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    assert lib.mrc_version() == 100


def test_host_tables_equal_oracle_tables():
    import mrc_oracle as o
    from mrcaudiocodec_b200.tables import Tables
    for sr, L in ((48000, 1024), (44100, 1024), (48000, 256), (48000, 2048)):
        t = Tables(L, sr)
        assert np.array_equal(t.kbd, o.window.transition_coeffs(L, L))
        assert np.array_equal(t.hann, o.window.hann_coeffs(2 * L))
        f = (np.arange(L) + 0.5) * ((float(sr) / L) / 2.)
        assert np.array_equal(t.bark, o.psychoac.Bark(f))
        assert np.array_equal(t.quiet, o.psychoac.Intensity(o.psychoac.Thresh(f)))
        assert np.array_equal(t.band_nlines, o.psychoac.AssignMDCTLinesFromFreqLimits(L, sr).astype(np.int32))
    t = Tables(1024, 48000)
    for i, T in enumerate(o.tables.TABLES):
        assert t.huff_escape[i] == T.escape
        for v in range(65):
            if v in T.codes:
                assert t.huff_len[i, v] == len(T.codes[v]) and t.huff_code[i, v] == int(T.codes[v], 2)
            else:
                assert t.huff_len[i, v] == 0


def test_block_switching_tables_equal_oracle_tables():
    """the three extra block geometries: window, Hann, Bark, threshold in quiet and the 9-band table
    as the oracle (= the reference) computes them for codingParams.a / .b"""
    import mrc_oracle as o
    from mrc_oracle.pacfile import CodingParams, sfbands_for
    from mrcaudiocodec_b200.tables import BlockTables, transient_sos
    for sr in (48000, 44100):
        for a, b in ((1024, 128), (128, 1024), (128, 128), (1024, 1024)):
            t = BlockTables(a, b, 1024, sr)
            half = (a + b) // 2
            assert np.array_equal(t.window, o.window.transition_coeffs(a, b))
            assert np.array_equal(t.hann, o.window.hann_coeffs(a + b))
            f = (np.arange(half) + 0.5) * ((float(sr) / half) / 2.)
            assert np.array_equal(t.bark, o.psychoac.Bark(f))
            assert np.array_equal(t.quiet, o.psychoac.Intensity(o.psychoac.Thresh(f)))
            cp = CodingParams()
            cp.a, cp.b, cp.nMDCTLines, cp.sampleRate = a, b, 1024, sr
            assert np.array_equal(t.band_nlines, sfbands_for(cp).nLines.astype(np.int32))
            assert t.n_bands == (25 if a == b == 1024 else 9)
        assert np.array_equal(transient_sos(sr), o.transient.design_sos(sr))


def test_no_cpu_fallback():
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mrcaudiocodec_b200 import Codec, _lib
    with pytest.raises(_lib.MrcError) as e:
        Codec()
    assert e.value.code == _lib.MRC_E_CUDA


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mrcaudiocodec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mrc_oracle" not in src and "ref_shim" not in src, f


def test_synth_is_deterministic():
    from mrcaudiocodec_b200 import synth
    a = synth.synth_clip(3, 12.5)
    b = synth.synth_clip(3, 12.5, threads=4)
    assert a.shape == (600000, 2) and np.array_equal(a, b)
    assert a.min() > -32768
    assert not np.any(a[3 * 48000:4 * 48000])            # the exact-silence second


def test_pcm_division_by_reciprocal_is_exact():
    """The kernels convert int16 PCM with x = 2|c| / 65535 (pcmfile.py:87-101, quantize.py:90-111) evaluated as
    q * rcp with one FMA correction step (mrc_math.cuh pcm_to_fraction).  Exhaustively: for all 32768 magnitudes the
    result equals the correctly rounded quotient, and the residual FMA is exact."""
    from fractions import Fraction as F
    rcp = float(F(1, 65535))
    assert rcp == 1.0 / 65535.0
    for c in range(32768):
        q = 2.0 * c
        y0 = float(F(q) * F(rcp))
        r_exact = F(q) - F(65535) * F(y0)
        r = float(r_exact)
        assert F(r) == r_exact
        assert float(F(y0) + F(r) * F(rcp)) == q / 65535.0


def _fft_r4_pos(n, logn):
    """mrc_fft.cuh: fft_r4_pos -- digit reversal in radix 4 with one innermost radix-2 digit when logn is odd"""
    m2 = logn & ~1
    t = n & ((1 << m2) - 1)
    b = int(format(t, "0%db" % m2)[::-1], 2) if m2 else 0
    b = ((b & 0xaaaaaaaa) >> 1) | ((b & 0x55555555) << 1)
    return ((n >> m2) | (b << 1)) if (logn & 1) else b


def _fft_swz(e, elem):
    """mrc_fft.cuh: fft_swz<double> (16-byte elements) / fft_swz<float> (8-byte elements)"""
    if elem == 16:
        return e ^ (((e >> 3) & 1) * 3) ^ (((e >> 4) & 1) * 6)
    return e ^ (((e >> 4) & 1) * 5) ^ (((e >> 5) & 1) * 10)


def _fft_place_index(idx, logn):
    """mrc_fft.cuh: fft_place_index"""
    if logn < 9:
        return idx
    lane, hi = idx & 31, idx >> 5
    y, rest = (lane ^ hi) & 7, hi >> 3
    if logn == 9:
        return lane | ((rest & 1) << 5) | (y << 6)
    if logn == 10:
        return lane | ((rest & 1) << 5) | ((y & 1) << 6) | (((rest >> 1) & 1) << 7) | ((y >> 1) << 8)
    return lane | ((rest & 7) << 5) | (y << 8) | ((rest >> 3) << 11)


def test_fft_swizzle_and_placement_are_conflict_free():
    """The index maps of the analysis kernel's transforms (mrc_fft.cuh), restated here: the swizzle is a bijection that
    is linear over GF(2); in every access pattern of the transform the lanes that share a shared-memory wavefront -- 8 for
    16-byte elements (fp64), 16 for 8-byte ones (fp32) -- hit that many different bank slots; the input placement is a
    bijection whose quarter warps read 8 different slots and write 8 different slots."""
    for elem, group in ((16, 8), (8, 16)):
        swz = lambda e: _fft_swz(e, elem)
        for logn in (6, 7, 8, 9, 10, 11):
            n = 1 << logn
            assert sorted(swz(e) for e in range(n)) == list(range(n))
            assert all(swz(a ^ b) == swz(a) ^ swz(b) for a in range(0, n, 7) for b in (1, 2, 3, 5, 64, n >> 1))
            slots = lambda es: len({swz(e) % group for e in es})
            if logn & 1:        # radix-2 stage: butterfly q of a warp covers elements 2q, 2q + 1
                for leg in (0, 1):
                    for q0 in range(0, n // 2, group):
                        assert slots(2 * (q0 + l) + leg for l in range(group)) == group
            h = 2 if (logn & 1) else 1
            while 4 * h <= n:
                logh = h.bit_length() - 1
                for leg in range(4):
                    for b0 in range(0, n // 4, group):
                        es = []
                        for l in range(group):
                            bi = b0 + l
                            j = bi & (h - 1)
                            es.append((((bi >> logh) << (logh + 2)) + j) + leg * h)
                        assert slots(es) == group, (elem, logn, h, leg)
                h *= 4
    # placement (fp64: the lanes of a quarter warp read 8 consecutive inputs and write 8 different slots)
    for logn in (9, 10, 11):
        n = 1 << logn
        src = [_fft_place_index(i, logn) for i in range(n)]
        assert sorted(src) == list(range(n))
        for q0 in range(0, n, 8):
            reads = src[q0:q0 + 8]
            assert len({r % 8 for r in reads}) == 8
            writes = [_fft_swz(_fft_r4_pos(r, logn), 16) for r in reads]
            assert len({w % 8 for w in writes}) == 8
        for w0 in range(0, n, 32):      # a warp reads 32 consecutive inputs
            assert sorted(src[w0:w0 + 32]) == list(range(min(src[w0:w0 + 32]), min(src[w0:w0 + 32]) + 32)) or \
                len({s & 31 for s in src[w0:w0 + 32]}) == 32


def test_whole_file_entry_points_reject_non_stereo_clips():
    """A mono array must not be re-read as stereo frames (even / odd samples as L / R); empty clips of any shape pass."""
    from mrcaudiocodec_b200.codec import Codec
    pcm, off = Codec._concat([np.zeros((0, 2), np.int16), np.ones((5, 2), np.int16), np.zeros(0, np.int16)])
    assert pcm.shape == (5, 2) and off.tolist() == [0, 0, 5, 5]
    for bad in (np.ones(10, np.int16), np.ones((10, 1), np.int16), np.ones((2, 10), np.int16)):
        with pytest.raises(ValueError):
            Codec._concat([bad])


def test_pac_header_and_chunk_index_on_golden_files():
    """pacfile.parse_header / chunk_index (host bookkeeping the CLI's decode uses) on the reference's own files: fields
    as pacfileThem.py:592-613 writes them, chunk chain ends exactly at the end of the file, two chunks per block."""
    import glob
    from mrcaudiocodec_b200 import pacfile
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
    assert files
    seen = 0
    for f in files:
        g = np.load(f, allow_pickle=True)
        if "pac" not in g.files:
            continue
        blob = g["pac"].tobytes()
        h = pacfile.parse_header(blob)
        assert h["nChannels"] == 2 and h["nBands"] == len(h["nLines"]) and sum(h["nLines"]) == h["nMDCTLines"]
        idx = pacfile.chunk_index(blob)
        assert idx and idx[-1][0] + idx[-1][1] == len(blob) and len(idx) % 2 == 0
        assert pacfile.huff_table_ids(blob).shape == (len(idx),)
        seen += 1
    assert seen >= 5
