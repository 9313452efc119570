"""Full-size runs (BASELINE.json configs 2 and 4 shapes) checked through size-independent properties: determinism,
batch == single, chunk structure, header, encode->decode round-trip SNR, and byte-exactness of a sampled window
against the oracle (a stream prefix is independent of what follows it)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MINUTES = float(os.environ.get("MRC_FULLSIZE_MINUTES", "60"))


def _snr_db(ref, got):
    ref = ref.astype(np.float64)
    err = got.astype(np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum(err ** 2), 1e-30))


def test_one_hour_stream_properties():
    import mrc_oracle as o
    from mrcaudiocodec_b200 import Codec, synth, pacfile
    pcm = synth.synth_clip(0, 60 * MINUTES, threads=8, fast=True)
    c = Codec()
    blob = c.encode_clips([pcm])[0]
    blob2 = c.encode_clips([pcm])[0]
    assert blob == blob2, "encode is not deterministic"
    c2 = Codec(chain_tables=False)                      # serial walk without the tabulated fast path: same bytes
    assert c2.encode_clips([pcm])[0] == blob
    c2.close()
    h = pacfile.parse_header(blob)
    assert h["nMDCTLines"] == 1024 and h["nBands"] == 25 and h["sampleRate"] == 48000
    nblk = c.n_blocks(pcm.shape[0])
    idx = pacfile.chunk_index(blob)
    assert len(idx) == 2 * nblk
    kbps = 8 * len(blob) / (pcm.shape[0] / 48000.) / 1000.
    assert 200 < kbps < 270, kbps                       # 2 x 128 kb/s nominal, reservoir keeps it near the target
    # the first 40 blocks of the stream depend only on the first 40 blocks of PCM: byte-exact vs the oracle
    n = 40
    ob, _ = o.driver.encode_pcm(pcm[:n * 1024], joint=True)
    end = idx[2 * n][0] - 4
    o_idx = pacfile.chunk_index(ob)
    assert blob[h["headerBytes"]:end] == ob[h["headerBytes"]:o_idx[2 * n][0] - 4]
    dec = c.decode_clips([blob])[0]
    assert dec.shape[0] == nblk * 1024
    m = pcm.shape[0]
    loud = np.abs(pcm.astype(np.int32)).max(axis=1) > 64
    snr = _snr_db(pcm[loud], dec[:m][loud])
    assert snr > 10.0, snr
    c.close()


def test_decode_scouts_change_nothing(monkeypatch):
    """The host's walk over the <L nBytes> chain with scouts (worker threads that walk later byte ranges ahead of the
    parser, from guessed chunk starts) and without: same PCM on a valid stream; on a stream whose length prefixes are
    damaged inside a scouted range, the same error, or the same PCM where the damage still parses (a prefix changed so
    that the chain lands on a later chunk start)."""
    from mrcaudiocodec_b200 import Codec, synth, pacfile, _lib
    pcm = synth.synth_clip(3, 240.0, threads=8, fast=True)           # ~7.7 MB of .pac: three scouted ranges
    c = Codec()
    blob = c.encode_clips([pcm])[0]
    idx = pacfile.chunk_index(blob)

    def run(b, scouts):
        monkeypatch.setenv("MRC_DECODE_SCOUTS", str(scouts))
        try:
            return c.decode_clips([b])[0]
        except _lib.MrcError as e:
            return (e.code, str(e))

    ref = run(blob, 0)
    for k in (1, 3, 8):
        got = run(blob, k)
        assert isinstance(got, np.ndarray) and np.array_equal(got, ref)
    # also as one clip of a batch next to small ones (exact scouts at clip starts)
    monkeypatch.setenv("MRC_DECODE_SCOUTS", "4")
    small = c.encode_clips([pcm[:48000], pcm[48000:2 * 48000]])
    multi = c.decode_clips([small[0], blob, small[1], blob])
    assert np.array_equal(multi[1], ref) and np.array_equal(multi[3], ref)
    assert np.array_equal(multi[0], run(small[0], 0)) and np.array_equal(multi[2], run(small[1], 0))
    for frac in (0.3, 0.55, 0.9):
        off, n = idx[int(frac * len(idx)) | 1]                        # a pair's second chunk, deep inside the stream
        bad = bytearray(blob)
        bad[off - 4:off] = (n + 1 << 20).to_bytes(4, "little")        # prefix far too long
        a, b = run(bytes(bad), 0), run(bytes(bad), 6)
        assert isinstance(a, tuple) and a == b, (a, b)
        bad = bytearray(blob)
        bad[off - 4:off] = (n - 1).to_bytes(4, "little")              # one byte short: the chain falls between chunks
        a, b = run(bytes(bad), 0), run(bytes(bad), 6)
        assert type(a) is type(b)
        assert a == b if isinstance(a, tuple) else np.array_equal(a, b)
    c.close()


def test_batch_of_clips_equals_singles():
    from mrcaudiocodec_b200 import Codec, synth
    n_clips = int(os.environ.get("MRC_FULLSIZE_CLIPS", "96"))
    clips = [synth.synth_clip(100 + i, 30, fast=True) for i in range(8)]
    # the batch repeats 8 distinct 30 s clips; every copy must encode to the same bytes as the single encode
    c = Codec()
    singles = [c.encode_clips([x])[0] for x in clips]
    batch = c.encode_clips([clips[i % 8] for i in range(n_clips)])
    for i, b in enumerate(batch):
        assert b == singles[i % 8], i
    c.close()


def test_factorised_equals_sequential_at_scale():
    """Ten minutes of music-like material and ten of the bench signal: the factorised evaluation of the masker
    spreading (default) and the pair-by-pair evaluation in the reference's order give the same bytes."""
    from mrcaudiocodec_b200 import Codec, synth
    minutes = min(MINUTES, 10.0)
    clips = [np.concatenate([synth.synth_music(100 + i, 30.0) for i in range(int(2 * minutes))], axis=0),
             synth.synth_clip(7, 60 * minutes, threads=8, fast=True)]
    cf, cs = Codec(), Codec(spreading="sequential")
    bf = cf.encode_clips(clips)
    bs = cs.encode_clips(clips)
    assert [len(b) for b in bf] == [len(b) for b in bs]
    assert bf == bs
    cf.close()
    cs.close()


def test_many_tiny_clips():
    """2000 clips of 0..3000 frames in one call (clips straddle waves, many clips per wave): a sample of them equals
    the single-clip encode, every clip has the chunk structure its length implies, and the batch decodes."""
    from mrcaudiocodec_b200 import Codec, synth, pacfile
    rng = np.random.default_rng(3)
    base = synth.synth_clip(11, 40, fast=True)
    lens = rng.integers(0, 3001, size=2000)
    starts = rng.integers(0, base.shape[0] - 3001, size=2000)
    clips = [base[s:s + n] for s, n in zip(starts, lens)]
    c = Codec()
    blobs = c.encode_clips(clips)
    assert len(blobs) == 2000
    for i in range(0, 2000, 97):
        assert c.encode_clips([clips[i]])[0] == blobs[i], i
    for i in range(0, 2000, 13):
        assert len(pacfile.chunk_index(blobs[i])) == 2 * c.n_blocks(len(clips[i])), i
    dec = c.decode_clips(blobs)
    assert [d.shape[0] for d in dec] == [c.n_blocks(len(x)) * 1024 for x in clips]
    c.close()


def test_block_switching_at_scale():
    """Ten minutes of the bench stream plus castanet-like material, encoded with the reference's transient
    detector / look-ahead loop, several waves long.  Size-independent properties: the written
    blocks are exactly the detector's plan and their sizes chain (a of a block = b of the one before); a stream
    prefix is byte-exact against the oracle; clips in a batch equal their single encodes; the stream decodes to
    the plan's length, and short blocks do buy what they are for: less pre-echo energy ahead of the hits."""
    import mrc_oracle as o
    from mrcaudiocodec_b200 import Codec, synth, pacfile
    minutes = min(MINUTES, 10.0)
    perc = np.concatenate([synth.synth_percussive(200 + i, 30.0) for i in range(int(2 * minutes))], axis=0)
    stream = synth.synth_clip(3, 60 * minutes, threads=8, fast=True)
    c = Codec(block_switching=True)
    flags, plan = c.detect_transients([perc, stream])
    blobs = c.encode_clips([perc, stream])
    assert blobs[0] == c.encode_clips([perc])[0] and blobs[1] == c.encode_clips([stream])[0]
    for blob, ab, pcm in zip(blobs, plan, (perc, stream)):
        idx = pacfile.chunk_index(blob)
        assert len(idx) == 2 * len(ab)
        got = []
        for k in range(0, len(idx), 2):
            q = (blob[idx[k][0]] >> 2) & 3
            assert ((blob[idx[k + 1][0]] >> 2) & 3) == q
            got.append((128 if q & 2 else 1024, 128 if q & 1 else 1024))
        assert got == ab
        assert all(got[i][0] == got[i - 1][1] for i in range(1, len(got))) and got[0][0] == 1024
        assert sum(b for _, b in ab) == (c.n_blocks(pcm.shape[0])) * 1024
    n_short = sum(1 for a, b in plan[0] if b == 128)
    assert n_short > 8 * 100, n_short                     # the castanets do switch, a lot
    # prefix byte-exactness: the first 30 blocks of PCM decide the first blocks of the stream (one block of look-ahead)
    n = 30
    ob, _, geom, _ = o.driver.encode_pcm_switched(perc[:(n + 1) * 1024], sos=c.sos)
    k = 0
    frames = 0
    while frames < n * 1024:                              # written blocks covering the first n PCM blocks
        frames += geom[k][1]
        k += 1
    h = pacfile.parse_header(blobs[0])
    o_idx, g_idx = pacfile.chunk_index(ob), pacfile.chunk_index(blobs[0])
    assert blobs[0][h["headerBytes"]:g_idx[2 * k][0] - 4] == ob[h["headerBytes"]:o_idx[2 * k][0] - 4]
    dec = c.decode_clips(blobs)
    assert [d.shape[0] for d in dec] == [c.n_blocks(x.shape[0]) * 1024 for x in (perc, stream)]
    plain = Codec()
    dec_plain = plain.decode_clips(plain.encode_clips([perc]))[0]
    plain.close()
    # pre-echo: error energy in the 256 samples before each switched block's position, switched vs long-only
    pos, e_sw, e_pl = 0, 0.0, 0.0
    ref = perc.astype(np.float64)
    for a, b in plan[0][:-1]:
        if b == 128 and a == 1024 and pos >= 1024:
            s = slice(pos - 256, pos)
            e_sw += np.sum((dec[0][s] - ref[s]) ** 2)
            e_pl += np.sum((dec_plain[s] - ref[s]) ** 2)
        pos += b
    assert e_sw < e_pl, (e_sw, e_pl)
    c.close()


def test_parity_census_bench_stream():
    """The benchmarked stream (BASELINE configs[1]; MRC_FULLSIZE_MINUTES of it) against the oracle well beyond its
    prefix: windows at the boundaries of its silent / -70 dBFS seconds (every third 10 s segment) and at random
    starts, each re-encoded by the oracle from the GPU's reservoir at the window's start; chunk bytes and the
    reservoir after every block must agree (scripts/parity_census.py; the full census is committed under profiles/)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import parity_census as pc
    clips = pc.make_material("stream", 60 * MINUTES)
    r = pc.census(clips, random_windows=int(2 * MINUTES) + 8, window_blocks=16, boundary_blocks=12, boundary_step=3)
    assert r["chunks_compared"] >= 2 * 16 * (int(2 * MINUTES) + 8) * 0.9, r
    assert r["chunks_mismatching"] == 0 and r["reservoir_mismatches"] == 0, r


def test_parity_census_music_and_batch():
    """Same census on dense-masker music-like material (one stream) and on a batch of independent clips."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import parity_census as pc
    minutes = min(MINUTES, 10.0)
    r = pc.census(pc.make_material("music", 60 * minutes), random_windows=int(3 * minutes) + 4, window_blocks=12,
                  with_boundaries=False)
    assert r["chunks_mismatching"] == 0 and r["reservoir_mismatches"] == 0, r
    r = pc.census(pc.make_material("batch", 60 * minutes), random_windows=int(4 * minutes) + 4, window_blocks=10,
                  with_boundaries=False)
    assert r["chunks_mismatching"] == 0 and r["reservoir_mismatches"] == 0, r
