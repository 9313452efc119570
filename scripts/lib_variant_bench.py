"""Development aid: the bench stream through a VARIANT build of the library (libmrc_<tag>.so next to libmrc.so, e.g. the
analysis kernel compiled with another -DMRC_NEAR_LOUD_N), bytes compared with the shipped build's.
usage: python scripts/lib_variant_bench.py <tag> [seconds]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import _lib  # noqa: E402
tag = sys.argv[1]
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 1800.0
if tag != "base":
    _lib.LIB_PATH = os.path.join(ROOT, "mrcaudiocodec_b200", "libmrc_%s.so" % tag)
from mrcaudiocodec_b200 import Codec, synth  # noqa: E402
import torch  # noqa: E402

pcm = synth.synth_clip(0, seconds, threads=8, fast=True)
c = Codec()
d_pcm = torch.from_numpy(pcm).cuda()
cap = int(2.0 * (128000.0 / 48000.0) * 2 * pcm.shape[0] / 8) + (1 << 20)
d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
off = np.array([0, pcm.shape[0]], dtype=np.int64)
for _ in range(2):
    boff = c.encode_batch_device(d_pcm.data_ptr(), off, d_out.data_ptr(), cap)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    boff = c.encode_batch_device(d_pcm.data_ptr(), off, d_out.data_ptr(), cap)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
t = c.last_timing()
import hashlib
h = hashlib.sha1(d_out[:int(boff[-1])].cpu().numpy().tobytes()).hexdigest()[:12]
print("%s: %.0f audio-s/s, analysis %.2f ms, general pairs %d, bytes %d sha1 %s" %
      (tag, seconds / dt, t["analysis_ms"], t["general_pairs"], int(boff[-1]), h))
c.close()
