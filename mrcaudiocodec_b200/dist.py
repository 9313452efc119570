"""Multi-GPU sharding of the encode/decode hot path: one process per GPU, torch.distributed for the
plumbing.  Clips (files) are independent streams, so a batch shards across ranks with NO data-path collective: rank r
encodes the contiguous clip range shard_range(n_clips, r, world) with its own libmrc context.  The only exchange is
one all-gather of the per-clip bitstream lengths, from which every rank derives the byte offset of each of its clips
in the concatenated output (what a writer of one big archive, or of per-clip files in a shared index, needs).

Nothing here touches the bitstream: the bytes of a clip do not depend on which rank encoded it
(tests/test_dist_gloo.py checks the bookkeeping with world_size 2 on the gloo backend; tests/test_gpu_parity.py checks
that a sharded batch equals the single-context batch on the GPU).

A single long stream shards by BLOCK RANGE (encode_stream_sharded): rank r encodes a contiguous range of the stream's
blocks.  A block inherits two things from the block before it -- the previous n_mdct_lines frames (pacfileThem.py
:799-802), which the shard reads as an N/2-sample halo from the PCM itself, and codingParams.bitReservoir
(codecThem.py:224,274,332,391,503), one int.  Everything that does not depend on the reservoir (transform,
psychoacoustics, bit prices, the composed reservoir maps) runs on all ranks at once; then the reservoir travels down the
ranks, one point-to-point message of one int32 per boundary (the serial pass over a shard's composed maps takes well
under a millisecond), and every rank packs its chunks as soon as it has handed the reservoir on.  The concatenated
shards are byte-identical to the single-GPU file.  Still no collective on the data path: the all-gather of the shards'
byte counts gives the offsets for the concatenation."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, balanced range [lo, hi) of items for `rank`: the first (n_items % world) ranks get one more."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_shard_range(n_blocks, rank, world, hop_blocks=360):
    """Block range [lo, hi) of `rank` when ONE stream is sharded over `world` ranks.  The reservoir reaches rank r one
    hand-off later than rank r - 1, so later ranks have that much longer for the work that does not depend on it: rank r
    gets hop_blocks more blocks than rank r - 1 (one hand-off, ~0.1 ms, is worth ~360 blocks of analysis on a B200), so
    that every rank is ready about when its reservoir arrives.  (The first rank is no exception: its serial pass takes as
    long as the pass the others run ahead of their reservoir.  Giving it fewer blocks was measured and only moved the wait
    to the last rank.)  Short streams (fewer than 4096 blocks per rank) are split evenly."""
    n_blocks, world = int(n_blocks), int(world)
    if world <= 1 or n_blocks < world * 4096 or hop_blocks <= 0:
        return shard_range(n_blocks, rank, world)
    base = (n_blocks - hop_blocks * world * (world - 1) // 2) // world
    if base < 8 * hop_blocks:                      # many ranks on a short stream: the skew would eat the first shards
        return shard_range(n_blocks, rank, world)
    sizes = [base + hop_blocks * r for r in range(world)]
    rest = n_blocks - sum(sizes)                   # 0 <= rest < world: the last ranks take one more
    for r in range(world - rest, world):
        sizes[r] += 1
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


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


def encode_stream_sharded(codec, pcm_shard, pcm_frame0, total_frames, group=None, device=None, out=None,
                          device_ptrs=None, relay_group=None):
    """One stream of total_frames frames encoded by all ranks, rank r taking the block range
    stream_shard_range(ceil(total_frames / L), r, world).  pcm_shard: int16 [frames, 2] starting at stream frame pcm_frame0 and
    covering codec.shard_pcm_range(total_frames, first_block, n_blocks) (each rank only needs its own range: generate or
    read per rank).  Returns (this rank's bytes, shard byte offsets int64 [world + 1]): rank r's bytes belong at
    offsets[r] of the .pac file.  relay_group: a process group with a CPU backend (gloo) for the hand-off of the reservoir
    -- the value is on the host on both sides anyway (the library's callbacks), so a CPU message saves the device round
    trip of an NCCL send/recv pair; default: point-to-point messages on `group`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nblk = (int(total_frames) + codec.L - 1) // codec.L
    lo, hi = stream_shard_range(nblk, rank, world)
    if relay_group is not None:
        box = torch.zeros(1, dtype=torch.int32)
        pg = relay_group
    else:
        box = torch.zeros(1, dtype=torch.int32, device=device)
        pg = group

    def recv():
        if rank == 0:
            return 0
        dist.recv(box, src=rank - 1, group=pg)
        return int(box.item())

    def send(r):
        if rank + 1 < world:
            box.fill_(int(r))
            dist.send(box, dst=rank + 1, group=pg)

    blob = codec.encode_shard(pcm_shard, pcm_frame0, total_frames, lo, hi - lo, rank == 0, rank == world - 1, recv, send,
                              out=out, device_ptrs=device_ptrs)
    n = blob if device_ptrs is not None else len(blob)      # device-resident runs return the byte count only
    _, offsets = gather_clip_offsets([n], world, group, device) if world > 1 else \
        (None, np.array([0, n], dtype=np.int64))
    return blob, offsets
