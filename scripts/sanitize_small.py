"""Small end-to-end exercise for compute-sanitizer: joint + independent encode, decode, per-block seam, L=256/2048."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import Codec, synth

pcm = synth.synth_short(3, 0.3)
for kw in (dict(), dict(joint=False), dict(n_mdct_lines=256), dict(n_mdct_lines=2048), dict(precision="fp32"),
           dict(spreading="sequential")):
    c = Codec(**kw)
    blobs = c.encode_clips([pcm, pcm[:1500], np.zeros((0, 2), np.int16)])
    dec = c.decode_clips(blobs)
    a = c.stage_analysis([pcm[:5000]])
    q = c.stage_alloc_quant([pcm[:5000]])
    print(kw, [len(b) for b in blobs], [d.shape for d in dec])
    c.close()
print("sanitize_small done")
