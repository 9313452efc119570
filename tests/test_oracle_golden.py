"""Pin the oracle restatement against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py via oracle/ref_shim.py): .pac bytes, every per-block integer, float taps, decoded PCM."""
import numpy as np
import pytest

import mrc_oracle as o
from conftest import GOLDEN_CASES


def _line_aligned(block, c, sf):
    t = block["huffTable"][c]
    out = np.zeros(int(sf.nLines.sum()), np.int32)
    i = 0
    for b in range(sf.nBands):
        if block["bitAlloc"][c][b]:
            for j in range(int(sf.nLines[b])):
                v = block["mantissa"][c][i]
                if isinstance(v, str):
                    p = v.split("/")
                    T = o.tables.TABLES[t]
                    v = int(p[1]) if p[0] == T.escape_code else T.rev[p[0]]
                out[sf.lowerLine[b] + j] = int(v)
                i += 1
    return out


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference(golden, name):
    g = golden(name)
    sr, joint, tbps = int(g["sampleRate"]), bool(g["joint"]), float(g["tbps"])
    blob, blocks = o.driver.encode_pcm(g["pcm"], joint=joint, trace=True, sampleRate=sr, targetBitsPerSample=tbps)
    assert blob == g["pac"].tobytes(), "oracle .pac differs from the reference's"
    sf = o.psychoac.ScaleFactorBands(g["nLines"])
    assert len(blocks) == g["reservoir"].shape[0]
    for i, b in enumerate(blocks):
        assert int(b["joint"]) == int(g["isJoint"][i])
        assert list(b["overallScale"]) == g["overallScale"][i][:len(b["overallScale"])].tolist()
        if b["joint"]:
            assert list(b["ms_switch"]) == g["ms_switch"][i].tolist()
        assert b["reservoir"] == int(g["reservoir"][i])
        for c in range(2):
            assert np.array_equal(np.asarray(b["bitAlloc"][c]), g["bitAlloc"][i, c])
            assert np.array_equal(np.asarray(b["scaleFactor"][c]), g["scaleFactor"][i, c])
            assert b["huffTable"][c] == int(g["huffTable"][i, c])
            assert np.array_equal(_line_aligned(b, c, sf), g["mantissa"][i, c])
    # float taps (first blocks): unscaled MDCT lines and SMRs as the reference's MDCT()/CalcSMRs() returned them
    nf = g["mdct"].shape[0]
    for i in range(nf):
        tap = blocks[i]["tap"]
        taps = [tap] if isinstance(tap, dict) else tap
        k = 0
        for t in taps:
            for c in range(len(t["lines"])):
                scale = blocks[i]["overallScale"][k]
                np.testing.assert_allclose(t["lines"][c] / (1 << scale), g["mdct"][i, k], rtol=0, atol=1e-15)
                np.testing.assert_allclose(t["smr"][c], g["smr"][i, k], rtol=0, atol=1e-9)
                k += 1


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_decoder_matches_reference_decoder(golden, name):
    g = golden(name)
    dec = o.driver.decode_pac(g["pac"].tobytes(), joint=bool(g["joint"]))
    assert dec.shape == g["decoded"].shape
    assert np.array_equal(dec, g["decoded"])
    # sanity: B+1 PCM blocks out; the first B reconstruct the input (lossy)
    n = g["pcm"].shape[0]
    err = dec[:n].astype(np.int64) - g["pcm"].astype(np.int64)
    assert np.sqrt(np.mean(err.astype(np.float64) ** 2)) < 0.05 * 32768
