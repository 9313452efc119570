"""Where does the bit reservoir live along a stream?  (the serial walk's table covers R_in in [-128, 640))
usage: python scripts/reservoir_hist.py [seconds]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import Codec, synth
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 1200.0
pcm = synth.synth_clip(0, seconds, threads=8, fast=True)
c = Codec()
r = c.stage_alloc_quant([pcm])["reservoir"].astype(np.int64)      # reservoir AFTER each block = R_in of the next
step = int(100 * 46.875)
for i in range(0, len(r), step):
    w = r[i:i + step]
    print("blocks %6d..: min %6d  p10 %6d  median %6d  p90 %6d  max %7d   below -128: %5.1f%%  above 640: %5.1f%%" %
          (i, w.min(), np.percentile(w, 10), np.median(w), np.percentile(w, 90), w.max(),
           100.0 * np.mean(w < -128), 100.0 * np.mean(w >= 640)))
c.close()
