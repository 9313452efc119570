"""GPU parity tests (run on the B200 box): libmrc.so through the C ABI vs (a) the golden vectors produced by the
unmodified reference and (b) the oracle restatement on fresh seeds.  Integer outputs and bytes must be identical in
fp64 mode; float taps are compared with the tolerance written next to each assert."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def codecs():
    from mrcaudiocodec_b200 import Codec
    cache = {}

    def get(sr=48000, joint=True, tbps=128000. / 48000., precision="fp64", L=1024, spreading="factorised"):
        key = (sr, joint, tbps, precision, L, spreading)
        if key not in cache:
            cache[key] = Codec(sample_rate=sr, joint=joint, target_bits_per_sample=tbps, precision=precision,
                               n_mdct_lines=L, spreading=spreading)
        return cache[key]
    yield get
    for c in cache.values():
        c.close()


def _cfg(g):
    return int(g["sampleRate"]), bool(g["joint"]), float(g["tbps"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_bytes_identical(golden, codecs, name):
    g = golden(name)
    sr, joint, tbps = _cfg(g)
    blob = codecs(sr, joint, tbps).encode_clips([g["pcm"]])[0]
    assert blob == g["pac"].tobytes()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_stage_taps(golden, codecs, name):
    g = golden(name)
    sr, joint, tbps = _cfg(g)
    c = codecs(sr, joint, tbps)
    a = c.stage_analysis([g["pcm"]])
    q = c.stage_alloc_quant([g["pcm"]])
    isj = g["isJoint"]
    nB = isj.shape[0]
    for i in range(nB):
        ns = 4 if isj[i] else 2
        assert np.array_equal(a["overallScale"][i, :ns], g["overallScale"][i, :ns])
        if isj[i]:
            assert np.array_equal(a["ms_switch"][i], g["ms_switch"][i])
    for k in ("bitAlloc", "scaleFactor", "mantissa", "huffTable", "reservoir"):
        assert np.array_equal(q[k], g[k]), k
    for i in range(g["mdct"].shape[0]):
        ns = 4 if isj[i] else 2
        ref = g["mdct"][i, :ns]
        # MDCT lines: <= 1e-12 of the block maximum (the reference's own twiddle phases carry ~1e-13)
        assert np.abs(a["mdct"][i, :ns] - ref).max() <= 1e-12 * np.abs(ref).max() + 1e-300
        # SMR: <= 1e-9 dB
        assert np.abs(a["smr"][i, :ns] - g["smr"][i, :ns]).max() <= 1e-9


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_decode_matches_reference_decoder(golden, codecs, name):
    g = golden(name)
    sr, joint, tbps = _cfg(g)
    pcm = codecs(sr, joint, tbps).decode_clips([g["pac"].tobytes()])[0]
    assert pcm.shape == g["decoded"].shape
    d = np.abs(pcm.astype(np.int64) - g["decoded"].astype(np.int64))
    assert d.max() <= 1                                   # north_star: decoded PCM within 1 LSB
    assert np.count_nonzero(d) <= 2, "fp64 decode should reproduce the reference decoder almost everywhere"


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sequential_spreading_mode(golden, codecs, name):
    """MRC_FLAG_SPREAD_SEQUENTIAL sums the maskers pair by pair in the reference's order; the default factorised
    evaluation is the same sum re-associated.  Both must give the reference's bytes, and their SMRs must agree
    to 1e-11 dB (the golden SMRs are matched to 1e-9 dB by both)."""
    g = golden(name)
    sr, joint, tbps = _cfg(g)
    cs = codecs(sr, joint, tbps, spreading="sequential")
    cf = codecs(sr, joint, tbps)
    assert cs.encode_clips([g["pcm"]])[0] == g["pac"].tobytes()
    a_s = cs.stage_analysis([g["pcm"]])
    a_f = cf.stage_analysis([g["pcm"]])
    assert np.array_equal(a_s["n_peaks"], a_f["n_peaks"])
    assert np.abs(a_s["smr"] - a_f["smr"]).max() <= 1e-11
    n = g["mdct"].shape[0]
    isj = g["isJoint"]
    for i in range(n):
        ns = 4 if isj[i] else 2
        assert np.abs(a_s["smr"][i, :ns] - g["smr"][i, :ns]).max() <= 1e-9


@pytest.mark.parametrize("joint", [True, False])
def test_fresh_seed_vs_oracle(codecs, joint):
    import mrc_oracle as o
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_short(21, 0.6)
    ob, _ = o.driver.encode_pcm(pcm, joint=joint)
    c = codecs(48000, joint)
    gb = c.encode_clips([pcm])[0]
    assert gb == ob
    od = o.driver.decode_pac(ob, joint=joint)
    gd = c.decode_clips([gb])[0]
    assert gd.shape == od.shape
    assert np.abs(gd.astype(np.int64) - od.astype(np.int64)).max() <= 1


def test_ragged_batch_and_edge_clips(codecs):
    """empty clip, 1 frame, exactly L, L+1, silence, full scale incl. -32768 (Q1), identical L/R, hard-panned."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import synth
    rng = np.random.default_rng(5)
    base = synth.synth_short(31, 0.25)
    full = rng.integers(-32768, 32768, size=(3000, 2)).astype(np.int16)
    full[10, 0] = -32768
    full[11, 1] = -32768
    mono = base[:4000].copy()
    mono[:, 1] = mono[:, 0]
    pan = base[:4000].copy()
    pan[:, 1] = 0
    clips = [np.zeros((0, 2), np.int16), base[:1], base[:1024], base[:1025], np.zeros((2500, 2), np.int16),
             full, mono, pan, base]
    c = codecs(48000, True)
    blobs = c.encode_clips(clips)
    for i, (clip, blob) in enumerate(zip(clips, blobs)):
        ob, _ = o.driver.encode_pcm(clip, joint=True)
        assert blob == ob, "clip %d" % i
        assert c.encode_clips([clip])[0] == blob, "batch vs single, clip %d" % i
    dec = c.decode_clips(blobs)
    for i, (clip, blob, d) in enumerate(zip(clips, blobs, dec)):
        od = o.driver.decode_pac(blob, joint=True)
        assert d.shape == od.shape, "clip %d" % i
        assert d.shape[0] == c.n_blocks(clip.shape[0]) * 1024
        if d.size:
            assert np.abs(d.astype(np.int64) - od.astype(np.int64)).max() <= 1, "clip %d" % i


@pytest.mark.parametrize("seg_blocks", [0, 3, 32])
def test_chain_table_fast_path_vs_oracle(monkeypatch, seg_blocks):
    """The single-stream fast path (tabulated reservoir maps, composed over segments of `seg_blocks` blocks; 0 = the
    per-block maps walked one by one; mrc_chain.cu) forced on for short clips: bytes equal the oracle's, for joint and
    independent channels, including silence (reservoir far outside the table: closed form) and the blocks after it
    (complete walk), and for several clips in one wave (segments that straddle two clips are not composed)."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import Codec, synth
    monkeypatch.setenv("MRC_CHAIN_TABLE_MIN_BLOCKS", "1")
    monkeypatch.setenv("MRC_CHAIN_SEGMENT_BLOCKS", str(seg_blocks))
    pcm = synth.synth_short(77, 0.8)
    clips = [pcm, synth.synth_short(78, 0.33), pcm[:5000], synth.synth_short(79, 0.5)]
    for joint in (True, False):
        c = Codec(joint=joint)
        blob = c.encode_clips([pcm])[0]
        assert c.last_timing()["launches"] >= (9 if seg_blocks else 8)     # table (+ segment) + chain were launched
        blobs = c.encode_clips(clips)
        st = c.stage_alloc_quant([pcm])
        c.close()
        ob, tr = o.driver.encode_pcm(pcm, joint=joint, trace=True)
        assert blob == ob, joint
        assert st["reservoir"].tolist() == [b["reservoir"] for b in tr]
        for x, b in zip(clips, blobs):
            assert b == o.driver.encode_pcm(x, joint=joint)[0], joint
    c = Codec(target_bits_per_sample=64000. / 48000.)
    ob, _ = o.driver.encode_pcm(pcm, joint=True, targetBitsPerSample=64000. / 48000.)
    assert c.encode_clips([pcm])[0] == ob
    c.close()


def test_music_like_material_vs_oracle(codecs):
    """dense loud tonal maskers (most peaks above 40 dB SPL): the loud-masker branch of the factorised spreading and
    the band-maximum search, against the oracle; joint and independent channels."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_music(5, 0.9)
    for joint in (True, False):
        ob, _ = o.driver.encode_pcm(pcm, joint=joint)
        assert codecs(48000, joint).encode_clips([pcm])[0] == ob, joint
        assert codecs(48000, joint, spreading="sequential").encode_clips([pcm])[0] == ob, joint
    od = o.driver.decode_pac(ob, joint=False)
    gd = codecs(48000, False).decode_clips([ob])[0]
    assert np.abs(gd.astype(np.int64) - od.astype(np.int64)).max() <= 1


def test_sharded_batch_equals_single_context(codecs):
    """multi-GPU sharding (mrcaudiocodec_b200/dist.py) emulated in one process: three 'ranks' encode their
    contiguous clip ranges with their own contexts; bytes and global offsets equal the single-context batch."""
    from mrcaudiocodec_b200 import Codec, synth, dist as mdist
    clips = [synth.synth_short(60 + i, 0.1 + 0.05 * (i % 4)) for i in range(10)]
    whole = codecs(48000, True).encode_clips(clips)
    world = 3
    sizes = []
    for r in range(world):
        lo, hi = mdist.shard_range(len(clips), r, world)
        c = Codec()
        blobs = c.encode_clips(clips[lo:hi])
        c.close()
        assert blobs == whole[lo:hi], r
        sizes += [len(b) for b in blobs]
    assert sizes == [len(b) for b in whole]


def test_reference_seam_drop_in(golden):
    """The reference-style PACFile loop of the oracle, with its `codec` module swapped for codec_gpu
    (the monkey-patch INTEGRATION.md describes), writes the reference's bytes."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import codec_gpu
    for name in ("joint48k_64", "indep48k_128"):
        g = golden(name)
        saved = o.pacfile.codec
        o.pacfile.codec = codec_gpu
        try:
            blob, _ = o.driver.encode_pcm(g["pcm"][:12 * 1024], joint=bool(g["joint"]), sampleRate=int(g["sampleRate"]),
                                          targetBitsPerSample=float(g["tbps"]))
        finally:
            o.pacfile.codec = saved
        ref, _ = o.driver.encode_pcm(g["pcm"][:12 * 1024], joint=bool(g["joint"]), sampleRate=int(g["sampleRate"]),
                                     targetBitsPerSample=float(g["tbps"]))
        assert blob == ref


def test_reference_seam_decode(golden):
    import mrc_oracle as o
    from mrcaudiocodec_b200 import codec_gpu
    g = golden("joint48k_64")
    blob = g["pac"].tobytes()
    saved = o.pacfile.codec
    o.pacfile.codec = codec_gpu
    try:
        dec = o.driver.decode_pac(blob, joint=True)
    finally:
        o.pacfile.codec = saved
    d = np.abs(dec.astype(np.int64) - g["decoded"].astype(np.int64))
    assert d.max() <= 1


def test_malformed_stream_is_rejected(golden, codecs):
    from mrcaudiocodec_b200 import _lib
    g = golden("joint48k_128")
    c = codecs(48000, True)
    blob = bytearray(g["pac"].tobytes())
    with pytest.raises(_lib.MrcError) as e:
        c.decode_clips([bytes(blob[:len(blob) - 7])])         # truncated last chunk
    assert e.value.code == _lib.MRC_E_FORMAT
    bad = bytearray(blob)
    bad[0:4] = b'RIFF'
    with pytest.raises(_lib.MrcError) as e:
        c.decode_clips([bytes(bad)])
    assert e.value.code == _lib.MRC_E_FORMAT
    # a chunk whose bits run out: shrink the first chunk's payload but keep the chain consistent
    from mrcaudiocodec_b200 import pacfile
    off, n = pacfile.chunk_index(bytes(blob))[2]
    cut = bytes(blob[:off - 4]) + (8).to_bytes(4, "little") + bytes(blob[off:off + 8]) + bytes(blob[off + n:])
    with pytest.raises(_lib.MrcError) as e:
        c.decode_clips([cut])
    assert e.value.code == _lib.MRC_E_FORMAT


def test_fp32_fast_mode_tolerances(golden, codecs):
    """fp32 fast mode: MDCT lines within 1e-5 of the block maximum, SMRs within 1e-5 relative (of the 96 dB
    reference level), decoded PCM within 1 LSB of the fp64 decode, and the stream decodes with the oracle decoder."""
    import mrc_oracle as o
    g = golden("joint48k_128")
    c32 = codecs(48000, True, precision="fp32")
    c64 = codecs(48000, True)
    a = c32.stage_analysis([g["pcm"]])
    for i in range(g["mdct"].shape[0]):
        ref = g["mdct"][i]
        assert np.abs(a["mdct"][i] - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-30
        assert np.abs(a["smr"][i] - g["smr"][i]).max() <= 1e-5 * 96 * 10
    blob = c32.encode_clips([g["pcm"]])[0]
    od = o.driver.decode_pac(blob, joint=True)              # reference-algorithm decoder accepts the fp32 stream
    assert od.shape == g["decoded"].shape
    d32 = c32.decode_clips([g["pac"].tobytes()])[0]
    d64 = c64.decode_clips([g["pac"].tobytes()])[0]
    assert np.abs(d32.astype(np.int64) - d64.astype(np.int64)).max() <= 1


def test_block_size_sweep_vs_oracle(codecs):
    """config 5: N = 512, 1024, 4096 (nMDCTLines 256, 512, 2048) stay byte-exact against the oracle."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_short(41, 0.3)
    for L in (256, 512, 2048):
        ob, _ = o.driver.encode_pcm(pcm, joint=True, nMDCTLines=L)
        c = codecs(48000, True, L=L)
        assert c.encode_clips([pcm])[0] == ob, L
        od = o.driver.decode_pac(ob, joint=True)
        gd = c.decode_clips([ob])[0]
        assert np.abs(gd.astype(np.int64) - od.astype(np.int64)).max() <= 1, L


def test_bitrate_sweep_vs_oracle(codecs):
    import mrc_oracle as o
    from mrcaudiocodec_b200 import synth
    pcm = synth.synth_short(43, 0.25)
    for kbps in (64, 96, 192, 256):
        tbps = kbps * 1000. / 48000.
        ob, _ = o.driver.encode_pcm(pcm, joint=True, targetBitsPerSample=tbps)
        assert codecs(48000, True, tbps).encode_clips([pcm])[0] == ob, kbps


def test_cli_wav_roundtrip(tmp_path):
    """in.wav -> in.pac -> in_decoded.wav through the command line front end; the .pac equals the
    oracle's file for the same PCM and the decoded WAV equals the oracle decoder's PCM within 1 LSB."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import cli, synth
    pcm = synth.synth_short(91, 0.4, sample_rate=44100)
    wav = str(tmp_path / "x.wav")
    cli.write_wav(wav, 44100, pcm)
    cli.main(["roundtrip", wav, "--kbps", "96"])
    blob = open(str(tmp_path / "x.pac"), "rb").read()
    ob, _ = o.driver.encode_pcm(pcm, joint=True, sampleRate=44100, targetBitsPerSample=96000. / 44100.)
    assert blob == ob
    sr, dec = cli.read_wav(str(tmp_path / "x_decoded.wav"))
    od = o.driver.decode_pac(ob, joint=True)
    assert sr == 44100 and dec.shape == od.shape
    assert np.abs(dec.astype(np.int64) - od.astype(np.int64)).max() <= 1


def test_buffer_too_small_reports_sizes(codecs):
    """C ABI error behaviour (include/mrc.h): an output buffer that is too small gives MRC_E_NOSPACE, never a write
    past the buffer, and the offsets array holds the sizes needed; the retry with that size gives the same bytes.
    The clips span several waves so that the wave-by-wave copies to the host are exercised with a tiny capacity."""
    import ctypes as C
    from mrcaudiocodec_b200 import _lib, synth
    c = codecs()
    clips = [synth.synth_clip(60 + i, 40.0, fast=True) for i in range(12)]     # 12 x 1876 blocks: two waves
    pcm, off = c._concat(clips)
    ref_out, ref_off = c.encode_batch(pcm, off)
    total = int(ref_off[-1])
    for cap in (0, 1000, total // 3, total - 1):
        guard = np.full(cap + 64, 0xA5, np.uint8)
        boff = np.zeros(len(clips) + 1, np.int64)
        rc = c.lib.mrc_encode_batch(c._ctx, pcm.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(clips),
                                    guard.ctypes.data_as(C.c_void_p), cap, boff.ctypes.data_as(C.c_void_p))
        assert rc == _lib.MRC_E_NOSPACE
        assert np.array_equal(boff, ref_off)
        assert np.all(guard[cap:] == 0xA5), "wrote past the caller's capacity"
    out = np.empty(total, np.uint8)
    boff = np.zeros(len(clips) + 1, np.int64)
    rc = c.lib.mrc_encode_batch(c._ctx, pcm.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(clips),
                                out.ctypes.data_as(C.c_void_p), total, boff.ctypes.data_as(C.c_void_p))
    assert rc == 0 and np.array_equal(out, ref_out[:total])
    # decode side: too few frames
    foff = np.zeros(len(clips) + 1, np.int64)
    small = np.full((1000 + 16, 2), 0x5A5A, np.int16)
    rc = c.lib.mrc_decode_batch(c._ctx, out.ctypes.data_as(C.c_void_p), boff.ctypes.data_as(C.c_void_p), len(clips),
                                small.ctypes.data_as(C.c_void_p), 1000, foff.ctypes.data_as(C.c_void_p))
    assert rc == _lib.MRC_E_NOSPACE
    assert [int(v) for v in np.diff(foff)] == [c.n_blocks(x.shape[0]) * 1024 for x in clips]
    assert np.all(small == 0x5A5A)
    dec, foff2 = c.decode_batch(out, boff)
    assert np.array_equal(foff2, foff)
    singles = c.decode_clips([out[boff[i]:boff[i + 1]].tobytes() for i in (0, 5, 11)])
    for k, i in enumerate((0, 5, 11)):
        assert np.array_equal(dec[foff[i]:foff[i + 1]], singles[k])


@pytest.mark.parametrize("joint", [True, False])
def test_stream_sharded_by_block_range_equals_whole(joint, monkeypatch):
    """One stream cut into block ranges (mrc_encode_shard: halo of n_mdct_lines frames from the PCM, the reservoir handed
    from shard to shard as one int): the shards, encoded one after the other on this GPU, concatenate to the bytes of
    the whole-stream encode and of the oracle -- with the cuts inside silence (reservoir far outside the tabulated
    range), at a ragged tail, and with an empty shard."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import Codec, synth
    monkeypatch.setenv("MRC_CHAIN_TABLE_MIN_BLOCKS", "1")
    pcm = synth.synth_short(31, 1.1)[:-333]                 # ragged: the last block is zero padded
    L = 1024
    total = pcm.shape[0]
    nblk = (total + L - 1) // L
    c = Codec(joint=joint)
    whole = c.encode_clips([pcm])[0]
    assert whole == o.driver.encode_pcm(pcm, joint=joint)[0]
    sil = int(0.35 * total) // L + 3                        # a block inside the clip's silent stretch
    for cuts in ([0, nblk], [0, 7, nblk], [0, sil, sil + 2, nblk], [0, 1, 1, nblk - 1, nblk]):
        parts, r = [], [0]
        for i in range(len(cuts) - 1):
            b0, n = cuts[i], cuts[i + 1] - cuts[i]
            lo, hi = c.shard_pcm_range(total, b0, n)
            got = []
            parts.append(c.encode_shard(pcm[lo:hi], lo, total, b0, n, i == 0, i == len(cuts) - 2, lambda: r[0],
                                        got.append).tobytes())
            assert len(got) == 1
            r[0] = got[0]
        assert b"".join(parts) == whole, cuts
    c.close()


def test_sine_window_selector():
    """window="sine" (window.py:10-25): MDCT lines equal the oracle's MDCT of the sine-windowed block, and encode ->
    decode with the same window reconstructs (Princen-Bradley: TDAC cancels), unlike decoding with the other window."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import Codec, synth
    from mrcaudiocodec_b200.tables import sine_window
    assert np.array_equal(sine_window(2048), o.window.sine_coeffs(2048))
    pcm = synth.synth_short(5, 0.4)
    c = Codec(window="sine")
    a = c.stage_analysis([pcm])
    L = 1024
    x = np.concatenate((np.zeros((L, 2), np.int16), pcm, np.zeros((2 * L, 2), np.int16)))
    for b in (0, 3, 11):
        for ch in range(2):
            blk = o.pcm.pcm_to_fraction(x[b * L:(b + 2) * L, ch])
            ref = o.mdct.MDCT(o.window.SineWindow(blk), L, L)
            assert np.abs(a["mdct"][b, ch] - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-30)
    blob = c.encode_clips([pcm])[0]
    dec = c.decode_clips([blob])[0][:pcm.shape[0]].astype(np.float64)
    ck = Codec()
    dec_wrong = ck.decode_clips([blob])[0][:pcm.shape[0]].astype(np.float64)
    ref = pcm.astype(np.float64)
    loud = np.abs(ref).max(axis=1) > 64
    snr = 10 * np.log10(np.sum(ref[loud] ** 2) / np.sum((dec[loud] - ref[loud]) ** 2))
    snr_wrong = 10 * np.log10(np.sum(ref[loud] ** 2) / np.sum((dec_wrong[loud] - ref[loud]) ** 2))
    assert snr > 10.0 and snr > snr_wrong + 3.0, (snr, snr_wrong)
    c.close()
    ck.close()


def test_reference_seam_mono(golden):
    """nChannels = 1 through the non-joint seam (codecThem.py:216 loops over codingParams.nChannels): the oracle's
    PACFile loop over a mono stream with its codec swapped for codec_gpu writes the oracle's own bytes -- Encode,
    EncodeNoHuff and Decode -- including the Huffman books (quiet passage) and the reservoir carried block to block."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import codec_gpu
    g = golden("indep48k_64")
    mono = np.ascontiguousarray(g["pcm"][:14 * 1024, :1])
    kw = dict(joint=False, sampleRate=int(g["sampleRate"]), targetBitsPerSample=float(g["tbps"]))
    ref, _ = o.driver.encode_pcm(mono, **kw)
    saved = o.pacfile.codec
    o.pacfile.codec = codec_gpu
    try:
        blob, _ = o.driver.encode_pcm(mono, **kw)
        dec = o.driver.decode_pac(blob, joint=False)
    finally:
        o.pacfile.codec = saved
    assert blob == ref
    assert pacfile_nchannels(blob) == 1
    od = o.driver.decode_pac(ref, joint=False)
    assert dec.shape == od.shape and np.abs(dec.astype(np.int64) - od.astype(np.int64)).max() <= 1


def pacfile_nchannels(blob):
    from mrcaudiocodec_b200 import pacfile
    return pacfile.parse_header(blob)["nChannels"]
