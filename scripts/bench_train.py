"""Measurement for SURVEY.md 8 f3 (Huffman table training): corpus frequency table on the GPU vs the oracle loop.
Prints one JSON line.  usage: python scripts/bench_train.py [minutes]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from mrcaudiocodec_b200 import synth, train

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
clips = [synth.synth_music(300 + i, 30.0, sample_rate=44100) for i in range(int(2 * minutes))]
train.corpus_table(clips[:1])                                   # warm up (context, tables)
t0 = time.perf_counter()
table = train.corpus_table(clips)
codes, esc = train.create_tree(table)
dt = time.perf_counter() - t0
import mrc_oracle as o
small = [c[:44100] for c in clips[:2]]
t1 = time.perf_counter()
o.huffman_train.corpus_table(small)
dto = time.perf_counter() - t1
print(json.dumps({"metric": "training corpus audio-seconds/sec (EncodeNoHuff + calculateFrequencies + createTree)",
                  "value": 60.0 * minutes / dt, "unit": "audio-s/s", "corpus_s": 60.0 * minutes, "seconds": dt,
                  "cpu_baseline": {"value": 2.0 / dto, "unit": "audio-s/s", "cores": 1, "kind": "port", "sample": "2 x 1 s clips"},
                  "escape_value": int(esc), "n_codes": len(codes), "largest_value": int(len(table) - 1)}))
