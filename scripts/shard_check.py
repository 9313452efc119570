#!/usr/bin/env python
"""One stream sharded by block range over the ranks of a torchrun launch (NCCL): the concatenated shards must be the
bytes of the single-GPU encode of the same stream (rank 0 encodes the whole stream too and compares).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/shard_check.py [seconds]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from mrcaudiocodec_b200 import Codec, synth, dist as mdist
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    codec = Codec(device=local)
    L = codec.L
    total = int(round(seconds * 48000)) - 777                      # ragged tail
    nblk = (total + L - 1) // L
    lo, hi = mdist.stream_shard_range(nblk, rank, world)
    f0, f1 = codec.shard_pcm_range(total, lo, hi - lo)
    pcm = synth.synth_range(0, f0, f1, seconds, threads=4, fast=True)[:max(0, min(f1, total) - f0)]
    blob, offsets = mdist.encode_stream_sharded(codec, pcm, f0, total, device=dev)
    # gather the shards on rank 0 (test plumbing only: variable sizes, so pad to the largest)
    width = int(np.max(np.diff(offsets)))
    send = torch.zeros(width, dtype=torch.uint8, device=dev)
    send[:blob.size] = torch.from_numpy(np.ascontiguousarray(blob)).to(dev)
    recv = [torch.zeros(width, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0)
    if rank == 0:
        got = b"".join(recv[r][:int(offsets[r + 1] - offsets[r])].cpu().numpy().tobytes() for r in range(world))
        whole = synth.synth_clip(0, seconds, threads=8, fast=True)[:total]
        ref = codec.encode_clips([whole])[0]
        ok = got == ref
        print(json.dumps({"world": world, "seconds": seconds, "frames": total, "bytes": len(ref), "shard_bytes":
                          np.diff(offsets).tolist(), "identical_to_single_gpu": bool(ok)}))
        if not ok:
            raise SystemExit("sharded stream differs from the single-GPU encode")
    dist.barrier()
    codec.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
