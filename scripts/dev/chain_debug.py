"""dev: which segments does the serial pass walk through block by block?  (libmrc_dbg.so = -DMRC_DEBUG_CHAIN)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "mrcaudiocodec_b200", "libmrc_dbg.so")
from mrcaudiocodec_b200 import Codec, synth
pcm = synth.synth_clip(0, 400.0, threads=8, fast=True)
c = Codec()
st = c.stage_reservoir(pcm, [0, pcm.shape[0]])
r = st["reservoir"]
import numpy as np
print("R percentiles", np.percentile(r, [0, 1, 5, 25, 50, 75, 95, 99, 100]))
for b in (140, 141, 142, 150, 186, 187, 190, 191, 192, 281, 282):
    print(b, r[b - 3:b + 3].tolist())
c.close()
