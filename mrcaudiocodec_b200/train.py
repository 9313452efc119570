"""Huffman table training.  The data-parallel part -- encoding the corpus with EncodeNoHuff and
histogramming the mantissas the way huffman.calculateFrequencies does -- runs on the GPU (mrc_mantissa_histogram);
what is left is a dozen numbers, handled here on the host:

  huffman.py:73-110   createTree(table, numEntries): keep the numEntries most frequent values, fold everything else
                      into the next one (the escape value), build the tree by merging the two rarest entries, re-sorting
                      (stably) after every merge
  huffman.py:10-31    createCodesArray: left edge '0', right edge '1'
  huffman_training_script.py:39-42, :80-83   nScaleBits 3, nMantSizeBits 5, 2.27 bits/sample, 10 entries

The result has the shape of the shipped training_data/*_table.pkl: ({value: (code string, length)}, escape value)."""
import numpy as np

from .codec import Codec

TRAIN_SCALE_BITS, TRAIN_MANT_SIZE_BITS, TRAIN_BITS_PER_SAMPLE, TRAIN_ENTRIES = 3, 5, 2.27, 10
MAX_BLOCKS_PER_CALL = 16384


def corpus_table(clips, sample_rate=44100, n_mdct_lines=1024, n_scale_bits=TRAIN_SCALE_BITS,
                 n_mant_size_bits=TRAIN_MANT_SIZE_BITS, target_bits_per_sample=TRAIN_BITS_PER_SAMPLE, device=0):
    """The frequency table the script ends its corpus loop with: counts [0 .. largest value seen] as a numpy array.
    Clips are processed in order, in groups of at most 16384 blocks per library call."""
    c = Codec(sample_rate=sample_rate, n_mdct_lines=n_mdct_lines, n_scale_bits=n_scale_bits,
              n_mant_size_bits=n_mant_size_bits, target_bits_per_sample=target_bits_per_sample, joint=False,
              device=device)
    table = np.zeros(65536, np.int64)
    gmax = -1
    group, blocks = [], 0

    def flush():
        nonlocal table, gmax, group, blocks
        if not group:
            return
        hist, gmax, reset = c.mantissa_histogram(group, gmax)
        table = hist if reset else table + hist
        group, blocks = [], 0

    for clip in clips:
        nb = (clip.shape[0] + c.L - 1) // c.L
        if nb > MAX_BLOCKS_PER_CALL:
            raise ValueError("a training clip may hold at most %d blocks" % MAX_BLOCKS_PER_CALL)
        if blocks + nb > MAX_BLOCKS_PER_CALL:
            flush()
        group.append(clip)
        blocks += nb
    flush()
    c.close()
    return table[:gmax + 1]


def create_tree(counts, n_entries=TRAIN_ENTRIES):
    """counts[v] = occurrences of mantissa value v (v = 0 .. len-1).  Returns (codes, escape_value) with
    codes = {value: (bit string, length)} for the n_entries most frequent values and the escape value."""
    counts = [int(x) for x in counts]
    if len(counts) <= n_entries:
        raise ValueError("need more than %d distinct values (the reference indexes entry %d)" % (n_entries, n_entries))
    # most frequent first; equal counts keep ascending value order (a stable sort of items listed by value)
    order = sorted(range(len(counts)), key=lambda v: -counts[v])
    kept = [(v, counts[v]) for v in order[:n_entries + 1]]
    rest = sum(counts[v] for v in order[n_entries + 1:])
    escape_value = kept[n_entries][0]
    kept[n_entries] = (escape_value, kept[n_entries][1] + rest)
    # entries: (payload, weight), payload = value or a (left, right) pair; rarest first, stable
    work = sorted(kept, key=lambda e: e[1])
    while len(work) > 1:
        left, right = work[0], work[1]
        work = sorted([((left, right), left[1] + right[1])] + work[2:], key=lambda e: e[1])
    codes = {}

    def walk(node, path):
        left, right = node
        for child, bit in ((left, "0"), (right, "1")):
            if isinstance(child[0], tuple):
                walk(child[0], path + bit)
            else:
                codes[child[0]] = (path + bit, len(path) + 1)

    walk(work[0][0], "")
    return codes, escape_value


def train(clips, **kw):
    """corpus -> (codes, escape value): the content of a *_table.pkl."""
    n_entries = kw.pop("n_entries", TRAIN_ENTRIES)
    return create_tree(corpus_table(clips, **kw), n_entries)
