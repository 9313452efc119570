"""Multi-GPU sharding of the encode/decode hot path: one process per GPU, torch.distributed for the
plumbing.  Clips (files) are independent streams, so a batch shards across ranks with NO data-path collective: rank r
encodes the contiguous clip range shard_range(n_clips, r, world) with its own libmrc context.  The only exchange is
one all-gather of the per-clip bitstream lengths, from which every rank derives the byte offset of each of its clips
in the concatenated output (what a writer of one big archive, or of per-clip files in a shared index, needs).

Nothing here touches the bitstream: the bytes of a clip do not depend on which rank encoded it
(tests/test_dist_gloo.py checks the bookkeeping with world_size 2 on the gloo backend; tests/test_gpu_parity.py checks
that a sharded batch equals the single-context batch on the GPU).

A single long stream does not shard this way in exact mode: block-to-block the reference carries bitReservoir
(codecThem.py:224,274,332,503), so a shard would need the reservoir of the block before its first one.  Transform and
psychoacoustics of a block range only need an N/2-sample halo of PCM, and libmrc already pipelines them against the
serial reservoir walk inside one GPU (DESIGN.md "Waves"); across GPUs the walk stays serial, so long single streams
are kept on one GPU and batches are what scales."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, balanced range [lo, hi) of items for `rank`: the first (n_items % world) ranks get one more."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_clip_offsets(local_sizes, n_clips, group=None, device=None):
    """All-gather the per-clip byte counts of every rank's shard and return (global_sizes int64 [n_clips],
    global_offsets int64 [n_clips+1]).  The one collective of the path: 8 bytes per clip, off the hot loop.
    Works on any backend (nccl on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    local_sizes = np.ascontiguousarray(local_sizes, dtype=np.int64)
    lo, hi = shard_range(n_clips, rank, world)
    if local_sizes.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d clips, its shard is [%d, %d)" % (rank, local_sizes.shape[0], lo, hi))
    if world == 1:
        sizes = local_sizes.copy()
    else:
        width = max(shard_sizes(n_clips, world))               # equal-sized messages: pad the short shards
        send = torch.zeros(width, dtype=torch.int64, device=device)
        if hi > lo:
            send[:hi - lo] = torch.from_numpy(local_sizes).to(send.device)
        recv = [torch.zeros(width, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(recv, send, group=group)
        parts = []
        for r in range(world):
            a, b = shard_range(n_clips, r, world)
            parts.append(recv[r][:b - a].cpu().numpy())
        sizes = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    offsets = np.zeros(n_clips + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    return sizes, offsets


def encode_sharded(codec, clips, n_clips_global, group=None, device=None):
    """Encode this rank's shard (`clips`: the clips of shard_range(n_clips_global, rank, world), in order) and
    return (blobs, global_offsets): the .pac bytes of the local clips and the byte offset every clip of the whole
    batch has in the concatenated output."""
    blobs = codec.encode_clips(clips)
    _, offsets = gather_clip_offsets([len(b) for b in blobs], n_clips_global, group, device)
    return blobs, offsets


def write_concatenated(path, blobs, offsets, rank, world):
    """Every rank writes its slice of the concatenated output at its own offset (the file must exist with the final
    size; rank 0 creates it)."""
    n = len(offsets) - 1
    lo, hi = shard_range(n, rank, world)
    with open(path, "r+b") as fh:
        for i, b in zip(range(lo, hi), blobs):
            fh.seek(int(offsets[i]))
            fh.write(b)
