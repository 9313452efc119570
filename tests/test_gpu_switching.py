"""Block switching on the GPU: transient detector, short / transition block geometries, switched
streams and their decoding -- libmrc.so through the C ABI against the golden vectors the reference produced
(tests/golden/switched*.npz) and against the oracle on fresh seeds.  Integers and bytes identical in fp64 mode."""
import numpy as np
import pytest

from conftest import SWITCHED_CASES

pytestmark = pytest.mark.gpu

GEOS = [(1024, 1024), (1024, 128), (128, 128), (128, 1024)]


@pytest.fixture(scope="module")
def codecs():
    from mrcaudiocodec_b200 import Codec
    cache = {}

    def get(sr=48000, tbps=128000. / 48000., precision="fp64", sos=None, switching=True, joint=True):
        key = (sr, tbps, precision, None if sos is None else sos.tobytes(), switching, joint)
        if key not in cache:
            cache[key] = Codec(sample_rate=sr, joint=joint, target_bits_per_sample=tbps, precision=precision,
                               block_switching=switching, switch_tables=True, transient_sos_sections=sos)
        return cache[key]
    yield get
    for c in cache.values():
        c.close()


def _flags(dets):
    return np.array([(1 if np.any(d == 1) else 0) | (2 if np.any(d > 1) else 0) for d in dets], np.uint8)


@pytest.mark.parametrize("name", SWITCHED_CASES)
def test_golden_detector_and_block_layout(golden, codecs, name):
    g = golden(name)
    c = codecs(int(g["sampleRate"]), float(g["tbps"]), sos=g["sos"])
    flags, ab = c.detect_transients([g["pcm"]])
    assert np.array_equal(flags, g["flags"])
    assert np.array_equal(np.array(ab[0], np.int32), g["geom"])


@pytest.mark.parametrize("name", SWITCHED_CASES)
def test_golden_switched_bytes_identical(golden, codecs, name):
    g = golden(name)
    c = codecs(int(g["sampleRate"]), float(g["tbps"]), sos=g["sos"])
    blob = c.encode_clips([g["pcm"]])[0]
    assert blob == g["pac"].tobytes()


@pytest.mark.parametrize("name", SWITCHED_CASES)
def test_golden_switched_decode(golden, codecs, name):
    g = golden(name)
    # decoding needs the block geometries but not the detector: any context with switch tables will do
    c = codecs(int(g["sampleRate"]), float(g["tbps"]), switching=False)
    pcm = c.decode_clips([g["pac"].tobytes()])[0]
    assert pcm.shape == g["decoded"].shape
    d = np.abs(pcm.astype(np.int64) - g["decoded"].astype(np.int64))
    assert d.max() <= 1                                   # decoded PCM within 1 LSB of the reference decoder
    assert np.count_nonzero(d) <= 4


@pytest.mark.parametrize("a,b", GEOS)
@pytest.mark.parametrize("joint", [True, False])
def test_seam_every_geometry_vs_oracle(codecs, a, b, joint):
    """codec_gpu.JointEncode / Encode / JointDecode / Decode with codingParams.a, .b as the reference's loop sets
    them, against the oracle's codec on the same data: every integer identical, decoded samples to 1e-12."""
    from mrcaudiocodec_b200 import codec_gpu
    from mrc_oracle import codec as ocodec, driver
    from mrc_oracle.pacfile import sfbands_for
    rng = np.random.default_rng(100 * a + b + int(joint))
    for trial, (tbps, res0) in enumerate([(128000. / 48000., 0), (64000. / 48000., 37), (2.86, -120)]):
        n = a + b
        t = np.arange(n) / 48000.
        x = [0.3 * np.sin(2 * np.pi * 1234.5 * t + ch) + 0.1 * np.sin(2 * np.pi * 5000.3 * t) * (ch + 1) +
             0.01 * rng.standard_normal(n) for ch in range(2)]
        if trial == 2:
            x[0][n // 2:n // 2 + 16] += 0.5 * rng.standard_normal(16)

        def params():
            cp = driver.make_params(targetBitsPerSample=tbps)
            cp.a, cp.b = a, b
            cp.bitReservoir = res0
            cp.sfBands = sfbands_for(cp)
            return cp
        cpo, cpg = params(), params()
        if joint:
            ro = ocodec.JointEncode([v.copy() for v in x], cpo)
            rg = codec_gpu.JointEncode([v.copy() for v in x], cpg)
        else:
            ro = ocodec.Encode([v.copy() for v in x], cpo)
            rg = codec_gpu.Encode([v.copy() for v in x], cpg)
        assert cpo.bitReservoir == cpg.bitReservoir
        for k in range(len(ro)):
            for u, v in zip(ro[k], rg[k]):
                if isinstance(u, (list, np.ndarray)):
                    assert list(u) == list(v), (k, a, b)
                else:
                    assert u == v, (k, a, b)
        # decode the oracle's integers on both sides
        nl = n // 2

        def aligned(mant, ba, table):
            from mrc_oracle.tables import TABLES
            out = np.zeros(1024, np.int32)
            sf = cpo.sfBands
            i = 0
            for bd in range(sf.nBands):
                if ba[bd]:
                    for j in range(int(sf.nLines[bd])):
                        v = mant[i]
                        if isinstance(v, str):
                            p = v.split("/")
                            v = int(p[1]) if len(p) > 1 else TABLES[table].rev[p[0]]
                        out[sf.lowerLine[bd] + j] = int(v)
                        i += 1
            return out
        if joint:
            S, A, M, O, ms, H = ro
            mm = [aligned(M[ch], A[ch], H[ch]) for ch in range(2)]
            yo = ocodec.JointDecode(S, A, mm, O, cpo, ms)
            yg = codec_gpu.JointDecode(S, A, mm, O, cpg, ms)
            for ch in range(2):
                assert np.abs(np.asarray(yo[ch]) - yg[ch]).max() <= 1e-12
        else:
            S, A, M, O, H = ro
            for ch in range(2):
                m1 = aligned(M[ch], A[ch], H[ch])
                yo = ocodec.Decode(S[ch], A[ch], m1, O[ch], cpo)
                yg = codec_gpu.Decode(S[ch], A[ch], m1, O[ch], cpg)
                assert np.abs(np.asarray(yo) - yg).max() <= 1e-12
        assert nl == len(yg) // 2 if not joint else True


@pytest.mark.parametrize("seed,kind,kbps", [(21, "percussive", 128), (22, "percussive", 64), (23, "short", 192)])
def test_fresh_switched_streams_vs_oracle(codecs, seed, kind, kbps):
    from mrcaudiocodec_b200 import synth
    from mrc_oracle import driver
    pcm = synth.synth_percussive(seed, 0.7) if kind == "percussive" else synth.synth_short(seed, 0.7)
    tbps = kbps * 1000. / 48000.
    c = codecs(48000, tbps)
    blob_o, _, geom, det = driver.encode_pcm_switched(pcm, sos=c.sos, targetBitsPerSample=tbps)
    flags, ab = c.detect_transients([pcm])
    assert np.array_equal(flags, _flags(det))
    assert ab[0] == [tuple(int(v) for v in g) for g in geom]
    blob = c.encode_clips([pcm])[0]
    assert blob == blob_o
    dec = c.decode_clips([blob])[0]
    ref = driver.decode_pac(blob_o)
    assert dec.shape == ref.shape
    assert np.abs(dec.astype(np.int64) - ref.astype(np.int64)).max() <= 1


def test_switched_batch_ragged_and_empty(codecs):
    """several clips in one call: empty, shorter than a block, not a multiple of the block, transient in the last
    block (no look-ahead), transient in the first 128 samples of a clip"""
    from mrcaudiocodec_b200 import synth
    from mrc_oracle import driver
    c = codecs(48000, 128000. / 48000.)
    rng = np.random.default_rng(3)
    clips = [np.zeros((0, 2), np.int16), synth.synth_percussive(31, 0.25)[:700], synth.synth_percussive(32, 0.3),
             synth.synth_short(33, 0.2)]
    x = (0.01 * rng.standard_normal((5000, 2)))
    x[10:40] += 0.6 * rng.standard_normal((30, 2))            # first segment of the first block
    x[4700:4730, 0] += 0.7 * rng.standard_normal(30)          # last block of the clip
    clips.append(np.round(np.clip(x, -0.999, 0.999) * 32767).astype(np.int16))
    blobs = c.encode_clips(clips)
    for pcm, blob in zip(clips, blobs):
        blob_o = driver.encode_pcm_switched(pcm, sos=c.sos)[0]
        assert blob == blob_o
    decs = c.decode_clips(blobs)
    for blob, d in zip(blobs, decs):
        ref = driver.decode_pac(blob)
        assert d.shape == ref.shape
        if d.size:
            assert np.abs(d.astype(np.int64) - ref.astype(np.int64)).max() <= 1


def test_long_only_stream_same_with_and_without_switching(codecs):
    """material without transients: the switched encoder writes the plain long-block stream"""
    from mrcaudiocodec_b200 import synth, Codec
    pcm = synth.synth_music(5, 0.5)
    c = codecs(48000, 128000. / 48000.)
    flags, ab = c.detect_transients([pcm])
    plain = Codec(sample_rate=48000, joint=True)
    try:
        if not flags.any():
            assert c.encode_clips([pcm])[0] == plain.encode_clips([pcm])[0]
        else:       # still must agree with the oracle
            from mrc_oracle import driver
            assert c.encode_clips([pcm])[0] == driver.encode_pcm_switched(pcm, sos=c.sos)[0]
    finally:
        plain.close()


def test_fp32_switched_stream_decodes(codecs):
    """fast mode: no byte identity, but the stream must parse, chain its block sizes, and decode close to fp64's"""
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_percussive(41, 0.6)
    c64, c32 = codecs(48000, 128000. / 48000.), codecs(48000, 128000. / 48000., precision="fp32")
    b64, b32 = c64.encode_clips([pcm])[0], c32.encode_clips([pcm])[0]
    d64, d32 = c64.decode_clips([b64])[0], c64.decode_clips([b32])[0]
    assert d64.shape == d32.shape
    e64 = d64[:len(pcm)].astype(np.float64) - pcm
    e32 = d32[:len(pcm)].astype(np.float64) - pcm
    assert abs(np.sqrt(np.mean(e32 ** 2)) - np.sqrt(np.mean(e64 ** 2))) <= 0.05 * np.sqrt(np.mean(e64 ** 2)) + 1.0
    d32b = c32.decode_clips([b64])[0]                     # fp32 decoder on the fp64 stream: within 1 LSB
    assert np.abs(d32b.astype(np.int64) - d64.astype(np.int64)).max() <= 1


def test_plain_context_rejects_switched_stream(golden):
    from mrcaudiocodec_b200 import Codec, _lib
    g = golden("switched48k_128")
    c = Codec(sample_rate=48000, joint=True)
    try:
        with pytest.raises(_lib.MrcError) as e:
            c.decode_clips([g["pac"].tobytes()])
        assert e.value.code == _lib.MRC_E_FORMAT
    finally:
        c.close()


def test_cli_roundtrip_with_block_switching(tmp_path):
    """`cli roundtrip in.wav --block-switching` = the reference's `python pacfileThem.py in.wav` flow: the .pac
    equals the oracle's switched stream, the decoded WAV the oracle decoder's PCM within 1 LSB."""
    from mrcaudiocodec_b200 import cli, synth, tables
    from mrc_oracle import driver
    pcm = synth.synth_percussive(77, 0.5)
    wav = str(tmp_path / "p.wav")
    cli.write_wav(wav, 48000, pcm)
    cli.main(["roundtrip", wav, "--block-switching"])
    blob = open(str(tmp_path / "p.pac"), "rb").read()
    ob = driver.encode_pcm_switched(pcm, sos=tables.transient_sos(48000))[0]
    assert blob == ob
    sr, dec = cli.read_wav(str(tmp_path / "p_decoded.wav"))
    od = driver.decode_pac(ob)
    assert sr == 48000 and dec.shape == od.shape
    assert np.abs(dec.astype(np.int64) - od.astype(np.int64)).max() <= 1


def test_switched_stream_at_training_parameters():
    """nScaleBits 3 / nMantSizeBits 5 / 2.27 bits per sample (huffman_training_script.py:39-42) with block switching,
    at 44.1 kHz: other header widths, other budgets, other band tables for all four block geometries."""
    from mrcaudiocodec_b200 import Codec, synth
    from mrc_oracle import driver
    pcm = synth.synth_percussive(51, 0.5, sample_rate=44100)
    c = Codec(sample_rate=44100, n_scale_bits=3, n_mant_size_bits=5, target_bits_per_sample=2.27, joint=True,
              block_switching=True)
    try:
        blob = c.encode_clips([pcm])[0]
        blob_o, _, geom, _ = driver.encode_pcm_switched(pcm, sos=c.sos, sampleRate=44100, nScaleBits=3, nMantSizeBits=5,
                                                        targetBitsPerSample=2.27)
        assert any(b == 128 for _, b in geom)
        assert blob == blob_o
        dec = c.decode_clips([blob])[0]
        ref = driver.decode_pac(blob_o)
        assert dec.shape == ref.shape and np.abs(dec.astype(np.int64) - ref.astype(np.int64)).max() <= 1
    finally:
        c.close()
