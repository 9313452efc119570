"""Block switching, CPU side: the oracle's transient detector and look-ahead driver against the
golden vectors that oracle/make_golden.py produced with the reference's own TransientDetector /
JointWriteDataBlock / Close (tests/golden/switched*.npz), and -- where /root/reference exists -- live."""
import os
import struct

import numpy as np
import pytest

from conftest import SWITCHED_CASES
from mrc_oracle import driver, transient
from mrc_oracle.pacfile import CodingParams


def _pairs(blob):
    """byte ranges of the block pairs after the header"""
    nBands = struct.unpack('<L', blob[22:26])[0]
    pos = 26 + 2 * nBands
    out = []
    while pos < len(blob):
        s = pos
        for _ in range(2):
            pos += 4 + struct.unpack('<L', blob[pos:pos + 4])[0]
        out.append((s, pos))
    return out


@pytest.mark.parametrize("name", SWITCHED_CASES)
def test_switched_stream_matches_reference(golden, name):
    g = golden(name)
    blob, _, geom, det = driver.encode_pcm_switched(g["pcm"], sos=g["sos"], sampleRate=int(g["sampleRate"]),
                                                   targetBitsPerSample=float(g["tbps"]))
    assert np.array_equal(np.array(geom, np.int32), g["geom"])
    flags = np.array([(1 if np.any(d == 1) else 0) | (2 if np.any(d > 1) else 0) for d in det], np.uint8)
    assert np.array_equal(flags, g["flags"])
    assert blob == g["pac"].tobytes()
    assert (g["geom"][:, 1] == 128).sum() >= 16          # the fixtures do exercise short blocks
    assert {tuple(r) for r in g["geom"].tolist()} == {(1024, 1024), (1024, 128), (128, 128), (128, 1024)}


@pytest.mark.parametrize("name", SWITCHED_CASES)
def test_shipped_loop_is_a_prefix(golden, name):
    """The shipped `__main__` never writes the last block it read (Q11); everything it writes before its flush block
    is byte-identical to the canonical stream."""
    g = golden(name)
    a, b = g["pac"].tobytes(), g["pac_shipped"].tobytes()
    pa, pb = _pairs(a), _pairs(b)
    assert len(pb) == len(g["geom_shipped"]) and len(pa) == len(g["geom"])
    n = len(pb) - 1
    assert n >= 1 and a[:pa[n - 1][1]] == b[:pb[n - 1][1]]


@pytest.mark.parametrize("name", SWITCHED_CASES)
def test_switched_decode_matches_reference_decoder(golden, name):
    g = golden(name)
    pcm = driver.decode_pac(g["pac"].tobytes())
    assert np.array_equal(pcm, g["decoded"])
    assert pcm.shape[0] == int(g["geom"][1:, 0].sum() + g["geom"][-1, 1])


def test_sosfilt_restatement_is_bit_identical():
    from scipy import signal
    rng = np.random.default_rng(5)
    for sr in (44100, 48000):
        sos = transient.design_sos(sr)
        assert sos.shape == (10, 6) and np.all(sos[:, 3] == 1.0)
        x = rng.uniform(-1, 1, 1024)
        assert np.array_equal(transient.sosfilt_plain(sos, x), signal.sosfilt(sos, x))


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs /root/reference")
def test_transient_detector_vs_reference_live():
    import ref_shim
    m = ref_shim.load()
    rng = np.random.default_rng(9)
    sos = transient.design_sos(48000)
    T = np.array([0.1, 0.075])

    def params():
        cp = CodingParams()
        cp.nChannels, cp.nSamplesPerBlock, cp.nSamplesShort = 2, 1024, 128
        cp.P = np.zeros((2, 9))
        return cp
    cpa, cpb = params(), params()
    for it in range(40):
        data = 0.02 * rng.standard_normal((2, 1024))
        for _ in range(int(rng.integers(0, 3))):
            s = int(rng.integers(0, 1000))
            data[int(rng.integers(0, 2)), s:s + 24] += rng.uniform(0.1, 0.9) * rng.standard_normal(min(24, 1024 - s))
        ra = transient.TransientDetector(data.copy(), cpa, sos, T)
        rb = m["pacfileThem"].TransientDetector(data.copy(), cpb, sos, T)
        assert np.array_equal(ra, rb) and np.array_equal(cpa.P, cpb.P)
