"""BASELINE.json configs[4]: bit-rate sweep 64..256 kb/s/ch at N = 2048 and block-size sweep N = 512..4096 (and N = 256
with the 9-band short table: the 25-band table has empty bands there and crashes the reference) -- the
roofline characterisation of the fused MDCT + psychoacoustics kernel.  One JSON line per configuration.
usage: python scripts/sweep.py [seconds]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import Codec, synth, tables  # noqa: E402
from bench import algorithmic_flops  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
pcm = synth.synth_clip(0, seconds, threads=8, fast=True)
off = np.array([0, pcm.shape[0]], dtype=np.int64)
peaks = None
configs = [(1024, k) for k in (64, 96, 128, 192, 256)] + [(L, 128) for L in (128, 256, 512, 2048)]
for L, kbps in configs:
    for precision in ("fp64", "fp32"):
        c = Codec(n_mdct_lines=L, target_bits_per_sample=kbps * 1000.0 / 48000.0, precision=precision,
                  band_limits=tables.SHORT_FREQ_LIMITS if L <= 128 else None)
        if peaks is None:
            peaks = c.measure_peaks()
        out = np.empty(int(2.5 * kbps * 1000 / 8 * 2 * seconds) + (1 << 22), dtype=np.uint8)
        for _ in range(2):
            c.encode_batch(pcm, off, out=out)
        ms, an, n = 0.0, 0.0, 3
        for _ in range(n):
            _, boff = c.encode_batch(pcm, off, out=out)
            t = c.last_timing()
            ms += t["total_ms"]
            an += t["analysis_ms"]
        nblk = c.n_blocks(pcm.shape[0])
        flops = algorithmic_flops(nblk - 1, 1, t["maskers"], L)
        peak = peaks["fp64_tflops"] if precision == "fp64" else peaks["fp32_tflops"]
        ach = flops / (an / n * 1e-3) / 1e12
        print(json.dumps({"config": {"n_mdct_lines": L, "N": 2 * L, "kbps_per_channel": kbps, "precision": precision,
                                     "n_bands": c.n_bands, "seconds": seconds},
                          "value": seconds / (ms / n * 1e-3), "unit": "audio-s/s (end to end, host buffers)",
                          "ms_per_step": ms / n, "analysis_ms": an / n, "blocks": nblk, "maskers": t["maskers"],
                          "bitstream_kbps": 8e-3 * int(boff[-1]) / seconds,
                          "roofline": {"bound": precision, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                                       "frac": ach / peak, "work": "SURVEY 8d reference formulation"}}))
        sys.stdout.flush()
        c.close()
