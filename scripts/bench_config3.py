"""BASELINE.json configs[3] at full size: a batch of 4096 independent 30 s clips (122 880 s of 48 kHz stereo audio,
23.6 GB of PCM) sharded contiguously over the ranks of one box, no data-path collective, one NCCL all-gather of the
per-clip bitstream lengths for the offsets of the concatenated output (mrcaudiocodec_b200.dist).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config3.py

The shard of a rank is built from 16 distinct synthetic clips (seeds rank*16 .. rank*16+15) repeated in order --
synthesising 512 different clips per rank would dominate the GPU-box time; every copy is still encoded on its own.
Timed end to end: PCM in pinned host memory -> .pac bytes in pinned host memory, H2D and D2H inside, barrier + device
synchronisation on both sides, max over ranks.  One JSON line from rank 0."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_CLIPS, CLIP_S, SR = int(os.environ.get("MRC_CONFIG3_CLIPS", "4096")), 30.0, 48000


def main():
    import torch
    import torch.distributed as dist
    from mrcaudiocodec_b200 import Codec, synth
    from mrcaudiocodec_b200 import dist as mdist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lo, hi = mdist.shard_range(N_CLIPS, rank, world)
    n_local = hi - lo
    distinct = [synth.synth_clip(rank * 16 + i, CLIP_S, fast=True) for i in range(16)]
    fr = distinct[0].shape[0]
    h_pcm = torch.empty((n_local * fr, 2), dtype=torch.int16).pin_memory()
    pcm = h_pcm.numpy()
    for i in range(n_local):
        pcm[i * fr:(i + 1) * fr] = distinct[i % 16]
    off = (np.arange(n_local + 1, dtype=np.int64) * fr)
    codec = Codec(device=local)
    cap = int(2.2 * 128000 / 8 * 2 * CLIP_S * n_local) + (1 << 22)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    out = h_out.numpy()

    def step():
        _, boff = codec.encode_batch(pcm, off, out=out)
        sizes, goff = mdist.gather_clip_offsets(np.diff(boff), N_CLIPS, device=dev)
        return boff, goff

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = 3, 2
    for _ in range(warmup):
        boff, goff = step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        boff, goff = step()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    # every copy of a distinct clip must have encoded to the same bytes (clips are independent streams)
    ok = all(out[boff[i]:boff[i + 1]].tobytes() == out[boff[i % 16]:boff[i % 16 + 1]].tobytes() for i in range(16, n_local, 37))
    t = codec.last_timing()
    if rank == 0:
        audio = N_CLIPS * CLIP_S * steps
        print(json.dumps({"metric": "encoded audio-seconds/sec, 48 kHz stereo 128 kb/s/ch", "value": audio / float(tt[0]),
                          "unit": "audio-s/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                          "ms_per_step": 1000.0 * float(tt[0]) / steps, "scaling": "strong", "dtype": "f64",
                          "data": "synthetic", "end_to_end": True,
                          "config": {"workload": "BASELINE configs[3]: %d independent clips of %.0f s sharded contiguously "
                                                 "over %d GPU(s) (%d clips, %.1f GB of PCM per rank), joint M/S, 128 kb/s/ch, "
                                                 "fp64; NCCL all-gather of per-clip bitstream lengths" %
                                                 (N_CLIPS, CLIP_S, world, n_local, n_local * fr * 4 / 1e9)},
                          "h2d_bytes_per_step_per_rank": int(n_local * fr * 4), "d2h_bytes_per_step_per_rank": int(boff[-1]),
                          "concatenated_bytes": int(goff[-1]), "copies_identical": bool(ok),
                          "rank0_stage_ms": {k: t[k] for k in ("analysis_ms", "cost_ms", "chain_ms", "pack_ms", "h2d_ms", "d2h_ms")},
                          "blocks_per_rank": t["blocks"]}))
    codec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
