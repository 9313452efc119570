"""SURVEY.md 8 f3 -- Huffman table training.
CPU: the oracle restatement against the shimmed reference (huffman.py calculateFrequencies / createTree /
createCodesArray; skipped where /root/reference is absent) and the product's host-side tree builder against the oracle.
GPU: the corpus frequency table from mrc_mantissa_histogram against the oracle's EncodeNoHuff loop."""
import numpy as np
import pytest

import mrc_oracle as o


def _random_tables(rng, n):
    for _ in range(n):
        m = int(rng.integers(12, 70))
        c = rng.integers(0, 50, size=m)
        c[rng.integers(0, m, size=3)] = rng.integers(0, 5000, size=3)      # a few dominant values, many ties
        yield {i: int(c[i]) for i in range(m)}


def test_oracle_tree_vs_reference():
    ref_shim = pytest.importorskip("ref_shim")
    if not ref_shim.available():
        pytest.skip("/root/reference not present")
    ref = ref_shim.load()["huffman"]
    rng = np.random.default_rng(11)
    # calculateFrequencies incl. its reset quirk, over several calls on one table
    tr, to = dict(), dict()
    for _ in range(40):
        data = rng.integers(0, int(rng.integers(3, 90)), size=int(rng.integers(0, 60))).tolist()
        tr = ref.calculateFrequencies(tr, data)
        to = o.huffman_train.calculateFrequencies(to, data)
        assert tr == to
    for table in _random_tables(rng, 40):
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):                    # the reference prints its tables
            root_r, esc_r = ref.createTree(dict(table), 10)
        root_o, esc_o = o.huffman_train.createTree(dict(table), 10)
        assert esc_r == esc_o
        codes_r = root_r[0].createCodesArray(dict())
        codes_o = o.huffman_train.createCodesArray(root_o[0], dict())
        assert codes_r == codes_o


def test_product_tree_builder_vs_oracle():
    from mrcaudiocodec_b200 import train
    rng = np.random.default_rng(12)
    for table in _random_tables(rng, 60):
        root, esc = o.huffman_train.createTree(dict(table), 10)
        want = o.huffman_train.createCodesArray(root[0], dict())
        counts = [table[i] for i in range(len(table))]
        got, gesc = train.create_tree(counts, 10)
        assert gesc == esc and got == want
    # the four shipped books are complete prefix codes of this shape: 10 values + escape
    for T in o.tables.TABLES:
        assert len(T.codes) >= 11 and abs(sum(2.0 ** -len(c) for c in T.codes.values()) - 1.0) < 1e-12


@pytest.mark.gpu
def test_corpus_table_vs_oracle():
    """EncodeNoHuff at the training parameters (nScaleBits 3, nMantSizeBits 5, 2.27 bits/sample, 44.1 kHz,
    independent channels) and calculateFrequencies over a small corpus: identical tables, hence identical code books.
    The corpus is ordered so that a later clip brings a new record value (the reset quirk fires across calls)."""
    from mrcaudiocodec_b200 import synth, train
    clips = [synth.synth_short(201, 0.25, sample_rate=44100) // 8, synth.synth_short(202, 0.3, sample_rate=44100),
             synth.synth_music(203, 0.3, sample_rate=44100)]
    want = o.huffman_train.corpus_table(clips)
    got = train.corpus_table(clips)
    m = max(want)
    assert got.shape[0] == m + 1
    assert np.array_equal(got, np.array([want[i] for i in range(m + 1)], dtype=np.int64))
    codes_o, esc_o = o.huffman_train.train(clips)
    codes_g, esc_g = train.train(clips)
    assert esc_o == esc_g and codes_o == codes_g


@pytest.mark.gpu
def test_training_parameters_encode_vs_oracle():
    """the codec at the training parameters (3 scale bits, 5 allocation bits) writes the oracle's bytes"""
    from mrcaudiocodec_b200 import Codec, synth
    pcm = synth.synth_short(204, 0.3, sample_rate=44100)
    ob, _ = o.driver.encode_pcm(pcm, joint=False, sampleRate=44100, nScaleBits=3, nMantSizeBits=5, targetBitsPerSample=2.27)
    c = Codec(sample_rate=44100, n_scale_bits=3, n_mant_size_bits=5, target_bits_per_sample=2.27, joint=False)
    assert c.encode_clips([pcm])[0] == ob
    od = o.driver.decode_pac(ob, joint=False)
    gd = c.decode_clips([ob])[0]
    assert np.abs(gd.astype(np.int64) - od.astype(np.int64)).max() <= 1
    c.close()
