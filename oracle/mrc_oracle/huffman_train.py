"""Huffman table training (TEST INFRASTRUCTURE, like the rest of this package).  Follows
/root/reference/huffman.py: calculateFrequencies :56-71, createTree :73-110, HuffmanNode.createCodesArray :10-31,
and the corpus loop of /root/reference/huffman_training_script.py:33-66 (EncodeNoHuff per block, independent
channels, nScaleBits 3, nMantSizeBits 5, 2.27 bits per sample; createTree(freq_table, 10) :80)."""
import numpy as np

from . import codec
from .pcm import pcm_to_fraction
from .psychoac import AssignMDCTLinesFromFreqLimits, ScaleFactorBands
from .driver import make_params


class Node(object):
    """:6-9"""

    def __init__(self, left=None, right=None):
        self.left = left
        self.right = right


def calculateFrequencies(table, data):
    """:56-71.  Quirk kept: current_max is local to the call, so the first never-seen value of a call zeroes every
    key below it (never-seen values are always new maxima, because all smaller keys get created here)."""
    current_max = -1
    for i in range(len(data)):
        key = int(data[i])
        if key in table:
            table[key] = table.get(key) + 1
        else:
            for k in range(current_max + 1, key):
                table[k] = 0
            table[key] = 1
            current_max = key
    return table


def createTree(table, numEntries):
    """:73-110 (prints dropped).  Returns ((root, weight), escape_value)."""
    freq_sorted = sorted(table.items(), key=lambda kv: kv[1], reverse=True)
    cutoff = numEntries
    for k in range(cutoff + 1, len(freq_sorted)):
        freq_sorted[cutoff] = (freq_sorted[cutoff][0], freq_sorted[cutoff][1] + freq_sorted[k][1])
    escape_value = freq_sorted[cutoff][0]
    del freq_sorted[cutoff + 1:]
    sorted_table = sorted(freq_sorted, key=lambda kv: kv[1])
    while len(sorted_table) > 1:
        l, r = sorted_table[0], sorted_table[1]
        sorted_table[0] = (Node(l, r), l[1] + r[1])
        del sorted_table[1]
        sorted_table = sorted(sorted_table, key=lambda kv: kv[1])
    return sorted_table[0], escape_value


def createCodesArray(node, huff_table, path=''):
    """:10-31"""
    for child, bit in ((node.left, '0'), (node.right, '1')):
        if child is None:
            continue
        if isinstance(child[0], Node):
            createCodesArray(child[0], huff_table, path + bit)
        else:
            huff_table[child[0]] = (path + bit, len(path) + 1)
    return huff_table


def corpus_table(clips, sampleRate=44100, nMDCTLines=1024, nScaleBits=3, nMantSizeBits=5, targetBitsPerSample=2.27):
    """huffman_training_script.py:31-66: one frequency table over all files; per file the reservoir starts at 0 and the
    prior block at zeros; the flush block Close() writes is not counted."""
    freq_table = dict()
    for pcm in clips:
        pcm = np.asarray(pcm, dtype=np.int16)
        n, nCh = pcm.shape
        cp = make_params(sampleRate=sampleRate, nChannels=nCh, numSamples=n, nMDCTLines=nMDCTLines,
                         nScaleBits=nScaleBits, nMantSizeBits=nMantSizeBits, targetBitsPerSample=targetBitsPerSample)
        cp.sfBands = ScaleFactorBands(AssignMDCTLinesFromFreqLimits(nMDCTLines, cp.sampleRate))
        prior = [np.zeros(nMDCTLines) for _ in range(nCh)]
        for b in range((n + nMDCTLines - 1) // nMDCTLines):
            seg = pcm[b * nMDCTLines:(b + 1) * nMDCTLines]
            if seg.shape[0] < nMDCTLines:
                seg = np.concatenate((seg, np.zeros((nMDCTLines - seg.shape[0], nCh), dtype=np.int16)))
            data = [pcm_to_fraction(seg[:, c]) for c in range(nCh)]
            full = [np.concatenate((prior[c], data[c])) for c in range(nCh)]
            prior = data
            S, A, M, O, H = codec.EncodeNoHuff(full, cp)
            for iCh in range(len(M)):
                freq_table = calculateFrequencies(freq_table, M[iCh])
    return freq_table


def train(clips, numEntries=10, **kw):
    table = corpus_table(clips, **kw)
    root, escape_value = createTree(table, numEntries)
    return createCodesArray(root[0], dict()), escape_value
