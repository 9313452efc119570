"""Deterministic synthetic 16-bit PCM for tests and benchmarks (SURVEY.md §8d "Synthetic inputs").

Per clip (seed = clip index) and per 10 s segment: six sines per channel, log-uniform in [60, 14000] Hz and
nudged off the FFT bin centres, amplitudes uniform in [0.01, 0.2], a shared set of frequencies with an
L/R phase offset in [0, pi/2] (correlated stereo, so M/S triggers) plus one hard-panned tone; white noise of
sigma 0.01 everywhere (keeps every FFT bin above rounding noise); 64-sample sigma-0.4 noise bursts every 0.5 s;
inside every 10 s segment one second of exact digital silence and one second at about -70 dBFS (these make
the Huffman tables win).  Clipped to +-0.999 and rounded to int16 (never -32768).

Segments are independent (seeded by (seed, segment index)), so long streams are generated segment by
segment, optionally on several threads."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

SEG_SECONDS = 10


def _segment(seed, seg, sample_rate, n, quiet=True, dtype=np.float64):
    rng = np.random.default_rng([int(seed), int(seg)])
    t = (np.arange(n, dtype=dtype) + dtype(seg) * dtype(SEG_SECONDS * sample_rate)) / dtype(sample_rate)
    f = np.exp(rng.uniform(np.log(60.0), np.log(14000.0), 6)) + 0.37 * sample_rate / 2048.0
    amp = rng.uniform(0.01, 0.2, 6)
    ph = rng.uniform(0, 2 * np.pi, 6)
    dph = rng.uniform(0, np.pi / 2, 6)
    fp = float(np.exp(rng.uniform(np.log(200.0), np.log(8000.0)))) + 0.37 * sample_rate / 2048.0
    ap = rng.uniform(0.02, 0.1)
    L = np.zeros(n, dtype=dtype)
    R = np.zeros(n, dtype=dtype)
    two_pi = dtype(2 * np.pi)
    for k in range(6):
        w = two_pi * dtype(f[k]) * t
        L += dtype(amp[k]) * np.sin(w + dtype(ph[k]))
        R += dtype(amp[k]) * np.sin(w + dtype(ph[k] + dph[k]))
    L += dtype(ap) * np.sin(two_pi * dtype(fp) * t)
    x = np.stack([L, R], axis=1)
    x += (0.01 * rng.standard_normal((n, 2))).astype(dtype)
    half = sample_rate // 2
    for s in range(0, n, half):
        e = min(s + 64, n)
        x[s:e] += (0.4 * rng.standard_normal((e - s, 2))).astype(dtype)
    if quiet:
        # second 3 of the segment: exact zeros; second 6: about -70 dBFS
        a, b = 3 * sample_rate, min(4 * sample_rate, n)
        if a < n:
            x[a:b] = 0.0
        a, b = 6 * sample_rate, min(7 * sample_rate, n)
        if a < n:
            x[a:b] *= dtype(10 ** (-70 / 20.0) / 0.2)
    np.clip(x, -0.999, 0.999, out=x)
    return np.round(x * 32767.0).astype(np.int16)


def synth_clip(seed, seconds, sample_rate=48000, quiet=True, threads=1, fast=False):
    """int16 [n, 2].  fast=True synthesises in float32 (several times quicker, for hour-long bench inputs)."""
    n = int(round(seconds * sample_rate))
    seg_n = SEG_SECONDS * sample_rate
    nseg = (n + seg_n - 1) // seg_n
    dtype = np.float32 if fast else np.float64
    jobs = [(seed, s, sample_rate, min(seg_n, n - s * seg_n), quiet, dtype) for s in range(nseg)]
    if threads > 1 and nseg > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda a: _segment(*a), jobs))
    else:
        parts = [_segment(*a) for a in jobs]
    return np.concatenate(parts, axis=0) if parts else np.zeros((0, 2), np.int16)


def synth_range(seed, frame_lo, frame_hi, total_seconds, sample_rate=48000, quiet=True, threads=1, fast=False):
    """frames [frame_lo, frame_hi) of synth_clip(seed, total_seconds): only the 10 s segments that overlap the range are
    synthesised (what one rank of a stream sharded by block range needs)."""
    n = int(round(total_seconds * sample_rate))
    frame_hi = min(int(frame_hi), n)
    frame_lo = max(0, min(int(frame_lo), frame_hi))
    seg_n = SEG_SECONDS * sample_rate
    dtype = np.float32 if fast else np.float64
    s0, s1 = frame_lo // seg_n, (frame_hi + seg_n - 1) // seg_n
    jobs = [(seed, s, sample_rate, min(seg_n, n - s * seg_n), quiet, dtype) for s in range(s0, s1)]
    if threads > 1 and len(jobs) > 1:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda a: _segment(*a), jobs))
    else:
        parts = [_segment(*a) for a in jobs]
    if not parts:
        return np.zeros((0, 2), np.int16)
    x = np.concatenate(parts, axis=0)
    return np.ascontiguousarray(x[frame_lo - s0 * seg_n:frame_hi - s0 * seg_n])


def synth_short(seed, seconds, sample_rate=48000):
    """A short test clip that still contains every regime: tones+noise, a transient, 0.2 s of exact silence,
    0.2 s at -70 dBFS.  Used for fixtures the CPU oracle / reference must finish in seconds."""
    n = int(round(seconds * sample_rate))
    x = _segment(seed, 0, sample_rate, n, quiet=False).astype(np.float64) / 32767.0
    a = int(0.35 * n)
    b = min(a + int(0.2 * sample_rate), n)
    x[a:b] = 0.0
    a = int(0.7 * n)
    b = min(a + int(0.2 * sample_rate), n)
    x[a:b] *= 10 ** (-70 / 20.0) / 0.2
    return np.round(np.clip(x, -0.999, 0.999) * 32767.0).astype(np.int16)


def synth_music(seed, seconds, sample_rate=48000):
    """Music-like test material (dense loud tonal maskers, unlike synth_clip's sparse tones over a noise floor): a
    new chord every 0.4 s of three to five notes, each a fundamental with a dozen decaying harmonics and a slow
    vibrato, panned individually, under an exponentially decaying attack envelope, plus a little noise; peaks around
    -3 dBFS.  Exercises the loud-masker path of the spreading (levels above 40 dB SPL) on most of the spectrum."""
    rng = np.random.default_rng([int(seed), 777])
    n = int(round(seconds * sample_rate))
    t = np.arange(n, dtype=np.float64) / sample_rate
    x = np.zeros((n, 2), dtype=np.float64)
    seg = int(0.4 * sample_rate)
    for s0 in range(0, n, seg):
        s1 = min(s0 + seg, n)
        tt = t[s0:s1] - t[s0]
        env = np.exp(-3.0 * tt) * (1.0 - np.exp(-400.0 * tt))
        for _ in range(int(rng.integers(3, 6))):
            f0 = 55.0 * 2.0 ** (rng.integers(0, 48) / 12.0) * (1.0 + 0.003 * rng.standard_normal())
            pan = rng.uniform(0.1, 0.9)
            vib = 1.0 + 0.004 * np.sin(2 * np.pi * rng.uniform(4, 7) * tt)
            ph = 2 * np.pi * f0 * np.cumsum(vib) / sample_rate
            note = np.zeros_like(tt)
            for h in range(1, 13):
                if f0 * h > 0.45 * sample_rate:
                    break
                note += (0.5 ** (0.5 * (h - 1))) * np.sin(h * ph + rng.uniform(0, 2 * np.pi))
            note *= env * rng.uniform(0.08, 0.25)
            x[s0:s1, 0] += note * np.sqrt(1.0 - pan)
            x[s0:s1, 1] += note * np.sqrt(pan)
    x += 0.002 * rng.standard_normal((n, 2))
    np.clip(x, -0.999, 0.999, out=x)
    return np.round(x * 32767.0).astype(np.int16)


def synth_percussive(seed, seconds, sample_rate=48000, hits_per_second=5.0):
    """Castanet-like material for block switching: a quiet tonal bed (two sines, -35 dBFS noise floor) under sharp
    hits at random times -- 2 ms attacks of band-limited noise decaying over 5..30 ms, random level and panning --
    so that the transient detector (pacfileThem.py:1021-1056) fires in varying 128-sample segments, in one or both
    channels, sometimes in consecutive blocks."""
    rng = np.random.default_rng([int(seed), 4242])
    n = int(round(seconds * sample_rate))
    t = np.arange(n, dtype=np.float64) / sample_rate
    x = np.zeros((n, 2), dtype=np.float64)
    for k in range(2):
        f = float(np.exp(rng.uniform(np.log(100.0), np.log(3000.0))))
        x[:, 0] += 0.05 * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
        x[:, 1] += 0.05 * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
    x += 0.004 * rng.standard_normal((n, 2))
    nhits = int(rng.poisson(hits_per_second * seconds)) + 1
    for _ in range(nhits):
        s0 = int(rng.integers(0, max(n - 1, 1)))
        dur = int(rng.uniform(0.005, 0.03) * sample_rate)
        m = min(dur, n - s0)
        if m <= 0:
            continue
        tt = np.arange(m) / float(sample_rate)
        env = np.exp(-tt / (dur / (5.0 * sample_rate)))
        burst = rng.standard_normal(m) * env * rng.uniform(0.05, 0.7)
        pan = rng.uniform(0.0, 1.0)
        x[s0:s0 + m, 0] += burst * np.sqrt(1.0 - pan)
        x[s0:s0 + m, 1] += burst * np.sqrt(pan)
    np.clip(x, -0.999, 0.999, out=x)
    return np.round(x * 32767.0).astype(np.int16)
