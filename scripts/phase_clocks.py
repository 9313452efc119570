"""Development tool: where does an analysis CTA spend its time?  Builds nothing itself: expects
mrcaudiocodec_b200/libmrc_clk.so (make -C mrcaudiocodec_b200/csrc clk: the library with -DMRC_PHASE_CLOCKS), encodes
--seconds of the bench stream and prints the cycles between phase boundaries, summed over CTAs (thread 0's clock
right after the barrier that ends each phase).  usage: python scripts/phase_clocks.py [seconds]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import _lib  # noqa: E402
_lib.LIB_PATH = os.path.join(ROOT, "mrcaudiocodec_b200", "libmrc_clk.so")
from mrcaudiocodec_b200 import Codec, synth  # noqa: E402

NAMES = ["0 load PCM", "1 MDCT (window, FFT, post-twiddle)", "2 ms_switch + overall max", "3 scale + Hann FFT + intensities",
         "4 peak compaction", "5 masker tables + scans + cell table", "6 pass 1 (bounds)", "7 pass 2a", "8 pass 2b",
         "9 band SMR", "10 hand-off", "11 grant order: keys", "12 grant order: merge + tokens"]

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
pcm = synth.synth_clip(0, seconds, threads=8, fast=True)
c = Codec()
lib = c.lib
lib.mrc_debug_phase_clocks.argtypes = [C.c_void_p, C.c_int]
c.encode_clips([pcm])
buf = np.zeros(32, np.uint64)
lib.mrc_debug_phase_clocks(None, 1)
c.encode_clips([pcm])
lib.mrc_debug_phase_clocks(buf.ctypes.data_as(C.c_void_p), 0)
tot = float(buf[:16].sum() + buf[23])      # 16..22, 24..27 are sub-step clocks of warp 0's lane 0
nblk = c.n_blocks(pcm.shape[0])
print("blocks %d, cycles per CTA %.0f" % (nblk, tot / nblk))
for i, n in enumerate(NAMES):
    print("%-44s %6.2f%%  %8.0f cycles/CTA" % (n, 100.0 * buf[i] / tot, buf[i] / nblk))
for i, n in ((13, "  1a MDCT: window + pre-twiddle"), (14, "  1b MDCT: FFT (then 1 = post-twiddle)"), (15, "  3a scale / Hann window placement"),
             (23, "  3b Hann FFT (then 3 = intensities)")):
    print("%-44s %6.2f%%  %8.0f cycles/CTA" % (n, 100.0 * buf[i] / tot, buf[i] / nblk))
sub = ["16 masker_range", "17 tails + loud maskers (one pass of 10**x)", "18 plateau sum + quiet", "19 -", "20 butterfly sum",
       "21 two log10 + division", "22 (count)", "23 -", "24 pass 2: band's best bound (warp 0's group, per band)",
       "25 pass 2: first complete threshold", "26 pass 2: scan for further candidates (+ their thresholds)",
       "27 pass 2: warp 0 waiting at the barrier (per spectrum)"]
ncomp = float(buf[22])
print("complete(): %.1f evaluations per block, cycles each (lane 0's clock):" % (ncomp / nblk))
for i, n in enumerate(sub):
    if i in (6, 7):
        continue
    print("  %-60s %8.0f" % (n, buf[16 + i] / max(ncomp, 1.0)))
bc = np.zeros((3, 32), np.uint64)
lib.mrc_debug_band_clocks.argtypes = [C.c_void_p]
lib.mrc_debug_band_clocks(bc.ctypes.data_as(C.c_void_p))
print('pass 2 per band (both encodes): tasks per block, cycles per task')
for b in range(25):
    if bc[1][b]:
        print('  band %2d  %5.2f  %7.0f' % (b, bc[1][b] / (2.0 * nblk), bc[0][b] / float(bc[1][b])))
c.close()
