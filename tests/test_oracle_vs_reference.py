"""Live comparison of the restatement with the shimmed reference (skipped where /root/reference is absent,
i.e. on the GPU box).  Function-level, so the closed forms in oracle/mrc_oracle/quantize.py etc. are checked
value for value over random inputs, not only through whole-file fixtures."""
import os

import numpy as np
import pytest

import mrc_oracle as o

ref_shim = pytest.importorskip("ref_shim")
pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")]


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def test_windows_bit_identical(ref):
    x = np.ones(2048)
    assert np.array_equal(ref["window"].KBDWindow(x), o.window.kbd_coeffs(2048))
    assert np.array_equal(ref["window"].HanningWindow(x), o.window.hann_coeffs(2048))
    assert np.array_equal(ref["window"].SineWindow(x), o.window.sine_coeffs(2048))
    assert np.array_equal(ref["window"].TransitionWindow(x, 1024, 1024), o.window.transition_coeffs(1024, 1024))
    xs = np.ones(1024 + 128)
    assert np.array_equal(ref["window"].TransitionWindow(xs, 1024, 128), o.window.transition_coeffs(1024, 128))


def test_mdct_imdct_bit_identical(ref):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(2048)
    X = ref["mdct"].MDCT(x, 1024, 1024)
    assert np.array_equal(X, o.mdct.MDCT(x, 1024, 1024))
    assert np.array_equal(ref["mdct"].IMDCT(X, 1024, 1024), o.mdct.IMDCT(X, 1024, 1024))


def test_quantizer_closed_forms(ref):
    rng = np.random.default_rng(2)
    q = ref["quantize"]
    for _ in range(300):
        ba = int(rng.integers(2, 17))
        x = rng.standard_normal(37) * 10.0 ** rng.uniform(-6, 0.2)
        x[rng.integers(0, 37)] = 0.0
        s_ref = q.ScaleFactor(np.max(np.abs(x)), 4, ba)
        assert s_ref == o.quantize.ScaleFactor(np.max(np.abs(x)), 4, ba)
        m_ref = q.vMantissa(x, s_ref, 4, ba)
        m = o.quantize.vMantissa(x, s_ref, 4, ba)
        assert np.array_equal(m_ref, m)
        assert np.array_equal(q.vDequantize(s_ref, m_ref, 4, ba), o.quantize.vDequantize(s_ref, m, 4, ba))
    for nb in (16, 20, 31):
        x = rng.uniform(-1.2, 1.2, 500)
        assert np.array_equal(q.vQuantizeUniform(x, nb), o.quantize.vQuantizeUniform(x, nb))
        c = rng.integers(0, 2 ** nb, 500)
        assert np.array_equal(q.vDequantizeUniform(c, nb), o.quantize.vDequantizeUniform(c, nb))


def test_masked_threshold_and_smr_bit_identical(ref):
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_short(7, 0.1)
    x = o.pcm.pcm_to_fraction(pcm[:2048, 0])
    sfb = o.psychoac.ScaleFactorBands(o.psychoac.AssignMDCTLinesFromFreqLimits(1024, 48000))
    rsfb = ref["psychoac"].ScaleFactorBands(ref["psychoac"].AssignMDCTLinesFromFreqLimits(1024, 48000))
    X = o.mdct.MDCT(o.window.TransitionWindow(x, 1024, 1024), 1024, 1024)
    thr_ref = ref["psychoac"].getMaskedThreshold(x, X, 0, 48000, rsfb)
    assert np.array_equal(thr_ref, o.psychoac.getMaskedThreshold(x, X, 0, 48000, sfb))
    assert np.array_equal(ref["psychoac"].CalcSMRs(x, X, 0, 48000, rsfb), o.psychoac.CalcSMRs(x, X, 0, 48000, sfb))


def test_bitalloc_matches(ref):
    rng = np.random.default_rng(3)
    nLines = o.psychoac.AssignMDCTLinesFromFreqLimits(1024, 48000).astype(int)
    nl2 = np.append(nLines, nLines)
    for _ in range(50):
        smr = rng.uniform(-30, 60, 50)
        B = float(rng.uniform(-100, 9000))
        a1, r1 = ref["bitalloc"].BitAlloc(B, 16, 50, nl2, smr.copy())
        a2, r2 = o.bitalloc.BitAlloc(B, 16, 50, nl2, smr.copy())
        assert np.array_equal(a1, a2) and r1 == r2


def test_whole_file_live(ref):
    """a fresh seed not in tests/golden, both stereo modes, encode + decode, byte for byte."""
    import ref_driver
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_short(11, 0.3)
    for joint in (True, False):
        rb, _ = ref_driver.ref_encode(pcm, joint=joint, trace=False)
        ob, _ = o.driver.encode_pcm(pcm, joint=joint)
        assert rb == ob
        assert np.array_equal(ref_driver.ref_decode(rb, joint=joint), o.driver.decode_pac(ob, joint=joint))
