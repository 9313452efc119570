#!/usr/bin/env python
"""Full-stream parity census: how many channel chunks of a long GPU-encoded stream equal the oracle's?

The GPU encodes the whole stream (BASELINE.json configs[1]: the 1 h bench stream; or music-like material; or a batch
of clips, configs[3] shape) and reports, per block, codingParams.bitReservoir after the block.  A block's bytes depend
on exactly two inherited things -- the previous 1024 PCM frames (pacfileThem.py:799-802) and the reservoir the block
before left (codecThem.py:391, :503, :274) -- so any window of consecutive blocks can be re-encoded by the oracle on
the host from (PCM of the window + one block, GPU reservoir at its start) and compared byte for byte
(/root/reference/pacfileThem.py:793-972 driven block by block).  Windows: every boundary of the stream's silent and
-70 dBFS seconds (where the Huffman books switch on and the reservoir swings) plus uniformly random starts; the
oracle's reservoir after every block is compared with the GPU's as well, so a window also vouches for the reservoir
it hands to the next block.  Oracle use is as the checker only (test infrastructure).

  python scripts/parity_census.py --material stream --seconds 3600 --random-windows 800 --out profiles/r02_parity_census.json
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

SR, L = 48000, 1024


def _oracle_window(job):
    """worker: (window id, prior PCM, window PCM, reservoir in, joint, kw) -> oracle chunk bytes and reservoirs"""
    import mrc_oracle as o
    wid, prior, blocks, r_in, joint, kw = job
    chunks, res = o.driver.encode_window(prior, blocks, r_in, joint=joint, **kw)
    return wid, chunks, res


def make_material(material, seconds, seed=0):
    from mrcaudiocodec_b200 import synth
    if material == "stream":
        return [synth.synth_clip(seed, seconds, threads=8, fast=True)]
    if material == "music":
        return [np.concatenate([synth.synth_music(100 + seed + i, 30.0) for i in range(max(1, int(seconds // 30)))], axis=0)]
    if material == "batch":
        return [synth.synth_clip(1000 + seed + i, 30.0, fast=True) for i in range(max(1, int(seconds // 30)))]
    raise ValueError(material)


def pick_windows(n_blocks, random_windows, window_blocks, boundary_blocks, rng, with_boundaries, boundary_step=1):
    """(first block, number of blocks) of every window; first block >= 1 (block 0 has no GPU reservoir before it: its
    reservoir is 0 by definition and it is covered by the window starting at 0)."""
    last = n_blocks - 1                       # the Close() flush block is non-joint: compared separately
    wins = [(0, min(window_blocks, last))]
    if with_boundaries:
        t = 0
        while t * SR < last * L:
            for s in (3, 4, 6, 7):            # synth_clip: second 3 of every 10 s segment silent, second 6 at -70 dBFS
                b = int((t + s) * SR // L)
                b0 = max(b - boundary_blocks // 2, 0)
                if b0 + boundary_blocks <= last:
                    wins.append((b0, boundary_blocks))
            t += 10 * boundary_step
    for _ in range(random_windows):
        b0 = int(rng.integers(0, max(last - window_blocks, 1)))
        wins.append((b0, min(window_blocks, last - b0)))
    return [w for w in wins if w[1] > 0]


def census(clips, joint=True, random_windows=256, window_blocks=32, boundary_blocks=16, procs=None, precision="fp64",
           with_boundaries=True, boundary_step=1, seed=1, pool=None, codec_kw=None, oracle_kw=None):
    """Encodes `clips` in one GPU call, re-encodes the sampled windows with the oracle, compares.  Returns a dict."""
    from mrcaudiocodec_b200 import Codec
    codec_kw = dict(codec_kw or {})
    oracle_kw = dict(oracle_kw or {})
    rng = np.random.default_rng(seed)
    own_pool = pool is None
    if own_pool:                              # spawned workers: nothing of this process's CUDA state is inherited
        procs = procs or max(1, min(len(os.sched_getaffinity(0)), 64))
        pool = mp.get_context("spawn").Pool(procs)
    t0 = time.time()
    c = Codec(joint=joint, precision=precision, **codec_kw)
    pcm, off = c._concat(clips)
    out, boff = c.encode_batch(pcm, off)
    st = c.stage_reservoir(pcm, off)
    c.close()
    t_gpu = time.time() - t0
    hdr = 26 + 2 * c.n_bands
    jobs, meta = [], []
    g = 0
    for ci in range(len(clips)):
        frames = int(off[ci + 1] - off[ci])
        nblk = c.n_blocks(frames)
        cb = st["chunkBytes"][g:g + nblk].astype(np.int64)
        res = st["reservoir"][g:g + nblk]
        start = boff[ci] + hdr + np.concatenate(([0], np.cumsum(8 + cb[:, 0] + cb[:, 1])))   # byte offset of every block pair
        assert start[-1] == boff[ci + 1], "chunk sizes do not add up to the clip's .pac size"
        x = pcm[off[ci]:off[ci + 1]]
        pad = (nblk - 1) * L - frames
        if pad:
            x = np.concatenate((x, np.zeros((pad, 2), np.int16)))
        for (b0, n) in pick_windows(nblk, random_windows if len(clips) == 1 else max(1, random_windows // len(clips)),
                                    window_blocks, boundary_blocks, rng, with_boundaries and len(clips) == 1,
                                    boundary_step):
            prior = x[(b0 - 1) * L:b0 * L] if b0 > 0 else np.zeros((L, 2), np.int16)
            r_in = int(res[b0 - 1]) if b0 > 0 else 0
            jobs.append((len(jobs), prior, x[b0 * L:(b0 + n) * L], r_in, joint, oracle_kw))
            meta.append((ci, b0, n, start, res))
        g += nblk
    t1 = time.time()
    compared = mism = res_mism = 0
    first = None
    blocks_seen = set()
    for wid, chunks, ores in pool.imap_unordered(_oracle_window, jobs, chunksize=1):
        ci, b0, n, start, res = meta[wid]
        for i in range(n):
            gbytes = out[start[b0 + i]:start[b0 + i + 1]].tobytes()
            ob = chunks[i]
            if (ci, b0 + i) not in blocks_seen:
                blocks_seen.add((ci, b0 + i))
                compared += 2
                if gbytes != ob:
                    # count per channel chunk: split at the first chunk's <L prefix
                    n0 = int.from_bytes(ob[:4], "little")
                    g0 = int.from_bytes(gbytes[:4], "little")
                    bad = (gbytes[:4 + g0] != ob[:4 + n0]) + (gbytes[4 + g0:] != ob[4 + n0:])
                    mism += bad
                    if first is None or (ci, b0 + i) < (first["clip"], first["block"]):
                        first = {"clip": ci, "block": b0 + i, "window_start": b0, "gpu_bytes": len(gbytes),
                                 "oracle_bytes": len(ob)}
                if int(res[b0 + i]) != ores[i]:
                    res_mism += 1
    if own_pool:
        pool.close()
        pool.join()
    return {"chunks_compared": compared, "chunks_mismatching": int(mism), "first_mismatch": first,
            "reservoir_mismatches": res_mism, "windows": len(jobs), "blocks_compared": len(blocks_seen),
            "blocks_in_stream": int(g), "clips": len(clips), "joint": bool(joint), "precision": precision,
            "audio_seconds_compared": len(blocks_seen) * L / float(SR),
            "gpu_seconds": round(t_gpu, 2), "oracle_seconds": round(time.time() - t1, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--material", default="stream", choices=["stream", "music", "batch"])
    ap.add_argument("--seconds", type=float, default=3600.0)
    ap.add_argument("--random-windows", type=int, default=800)
    ap.add_argument("--window-blocks", type=int, default=32)
    ap.add_argument("--boundary-blocks", type=int, default=16)
    ap.add_argument("--independent", action="store_true", help="WriteDataBlock flow (joint = 0)")
    ap.add_argument("--precision", default="fp64")
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    procs = a.procs or max(1, min(len(os.sched_getaffinity(0)), 64))
    pool = mp.get_context("spawn").Pool(procs)
    clips = make_material(a.material, a.seconds)
    r = census(clips, joint=not a.independent, random_windows=a.random_windows, window_blocks=a.window_blocks,
               boundary_blocks=a.boundary_blocks, precision=a.precision, pool=pool)
    pool.close()
    pool.join()
    r.update({"material": a.material, "seconds": a.seconds, "host_procs": procs})
    s = json.dumps(r)
    print(s)
    if a.out:
        with open(a.out, "a") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
