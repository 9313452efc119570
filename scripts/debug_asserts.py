"""Memory-safety run with the assert build (make -C mrcaudiocodec_b200/csrc debug -> libmrc_debug.so, -DMRC_DEBUG_ASSERTS):
every code path with its own buffer arithmetic on small inputs -- joint / independent, all block sizes, fp32, sequential
spreading, block switching, the tabulated and composed reservoir maps, shards, the per-block seam, training histogram,
decode incl. malformed input.  A failed device assert prints its condition and traps (the CUDA call then fails)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.path.join(ROOT, "mrcaudiocodec_b200", "libmrc_debug.so")
os.environ["MRC_CHAIN_TABLE_MIN_BLOCKS"] = "1"
from mrcaudiocodec_b200 import Codec, synth  # noqa: E402

pcm = synth.synth_short(3, 0.6)
perc = synth.synth_percussive(4, 0.5)
for kw in (dict(), dict(joint=False), dict(n_mdct_lines=256), dict(n_mdct_lines=512), dict(n_mdct_lines=2048),
           dict(precision="fp32"), dict(spreading="sequential"), dict(chain_tables=False), dict(window="sine"),
           dict(block_switching=True), dict(target_bits_per_sample=64000. / 48000.)):
    for seg in ("0", "3", "32"):
        os.environ["MRC_CHAIN_SEGMENT_BLOCKS"] = seg
        c = Codec(**kw)
        clips = [pcm, perc, pcm[:1500], np.zeros((0, 2), np.int16)]
        blobs = c.encode_clips(clips)
        dec = c.decode_clips(blobs)
        if not kw.get("block_switching"):
            c.stage_analysis([pcm[:5000]])
            c.stage_alloc_quant([pcm[:5000]])
            L = c.L
            total = pcm.shape[0]
            nblk = (total + L - 1) // L
            parts, r = [], [0]
            for b0, n, first, last in ((0, 5, True, False), (5, nblk - 5, False, True)):
                lo, hi = c.shard_pcm_range(total, b0, n)
                parts.append(c.encode_shard(pcm[lo:hi], lo, total, b0, n, first, last, lambda: r[0],
                                            lambda v: r.__setitem__(0, v)).tobytes())
            assert b"".join(parts) == blobs[0], kw
        c.close()
    print("ok", kw, [len(b) for b in blobs], [d.shape[0] for d in dec])
c = Codec(joint=False, n_scale_bits=3, n_mant_size_bits=5, target_bits_per_sample=2.27)
c.mantissa_histogram([pcm])
c.close()
c = Codec()
r = 0
for b in range(3):
    x = np.zeros((2, 2048))
    x[:, :] = np.random.default_rng(b).standard_normal((2, 2048)) * 0.1
    out, r = c.encode_block(x, 1, r)
    c.decode_block(True, out["scaleFactor"], out["bitAlloc"], out["mantissa"], out["overallScale"], out["ms_switch"])
    out, r = c.encode_block(x[:1], 4, r)
blob = bytearray(c.encode_clips([pcm])[0])
for cut in (7, 100, 1000):
    try:
        c.decode_clips([bytes(blob[:len(blob) - cut])])
    except _lib.MrcError:
        pass
blob[200] ^= 0xff
try:
    c.decode_clips([bytes(blob)])
except _lib.MrcError:
    pass
c.close()
print("debug_asserts done: no device assert fired")
