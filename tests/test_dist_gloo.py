"""world_size-2 test of the multi-GPU bookkeeping on CPU (gloo): sharding of a clip batch and the one collective of the
path, the all-gather of per-clip bitstream lengths -> global byte offsets (mrcaudiocodec_b200/dist.py).  The encoder is
replaced by a stand-in that returns deterministic byte strings, because the product encoder needs a GPU; what is under
test is the host logic around it.  The same functions run over NCCL in bench.py / on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeCodec(object):
    """stands in for Codec.encode_clips: clip -> bytes whose length and content depend only on the clip"""

    def encode_clips(self, clips):
        out = []
        for c in clips:
            n = 76 + 3 * int(c.shape[0] % 1000) + int(c[0, 0]) % 7
            out.append(bytes([int(c[0, 0]) % 251]) * n)
        return out


def _make_clips(n):
    rng = np.random.default_rng(7)
    return [rng.integers(-3000, 3000, size=(int(rng.integers(1, 5000)), 2)).astype(np.int16) for _ in range(n)]


def _worker(rank, world, port, n_clips, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mrcaudiocodec_b200 import dist as mdist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    clips = _make_clips(n_clips)
    lo, hi = mdist.shard_range(n_clips, rank, world)
    blobs, offsets = mdist.encode_sharded(_FakeCodec(), clips[lo:hi], n_clips)
    path = os.path.join(tmp, "all.pac")
    if rank == 0:
        with open(path, "wb") as fh:
            fh.truncate(int(offsets[-1]))
    dist.barrier()
    mdist.write_concatenated(path, blobs, offsets, rank, world)
    dist.barrier()
    np.save(os.path.join(tmp, "off%d.npy" % rank), offsets)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [7, 2, 1])
def test_sharded_encode_bookkeeping_world2(tmp_path, n_clips):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_clips, str(tmp_path)), nprocs=2, join=True)
    ref = _FakeCodec().encode_clips(_make_clips(n_clips))
    want = b"".join(ref)
    got = open(os.path.join(str(tmp_path), "all.pac"), "rb").read()
    assert got == want
    o0 = np.load(os.path.join(str(tmp_path), "off0.npy"))
    o1 = np.load(os.path.join(str(tmp_path), "off1.npy"))
    assert np.array_equal(o0, o1)
    assert np.array_equal(np.diff(o0), [len(b) for b in ref])


def test_shard_range_partitions():
    from mrcaudiocodec_b200 import dist as mdist
    for n in (0, 1, 5, 8, 4096, 4097):
        for w in (1, 2, 3, 4, 8):
            r = [mdist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_stream_shard_range_partitions():
    """block ranges of one sharded stream: contiguous, complete, even for short streams, and for long ones every rank
    gets one hand-off's worth of blocks more than the rank before it"""
    from mrcaudiocodec_b200 import dist as mdist
    for n in (0, 3, 4095, 8192, 32767, 32768, 33000, 168751, 1 << 20):
        for w in (1, 2, 3, 8, 64):
            r = [mdist.stream_shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert min(sizes) >= 0
            if w == 1 or n < w * 4096 or max(sizes) - min(sizes) <= 1:
                assert max(sizes) - min(sizes) <= 1
            else:
                assert all(0 <= sizes[i + 1] - sizes[i] - 360 <= 1 for i in range(w - 1)), sizes


class _FakeShardCodec(object):
    """stands in for Codec.encode_shard: the "reservoir" a shard hands on is a running checksum of the block indices it
    was given, so the test sees both the block ranges and the order of the hand-offs"""
    L = 1024

    def shard_pcm_range(self, total_frames, first_block, n_blocks):
        lo = max(first_block - 1, 0) * self.L
        return lo, max(min((first_block + n_blocks) * self.L, total_frames), lo)

    def encode_shard(self, pcm, pcm_frame0, total_frames, first_block, n_blocks, is_first, is_last, recv, send,
                     out=None, device_ptrs=None):
        lo, hi = self.shard_pcm_range(total_frames, first_block, n_blocks)
        assert pcm_frame0 <= lo and pcm_frame0 + pcm.shape[0] >= hi
        r = recv()
        assert (r == 0) == (first_block == 0)
        for b in range(first_block, first_block + n_blocks):
            r = (r * 31 + b + 1) % 1000003
        send(r)
        head = b"H" if is_first else b""
        tail = b"T" if is_last else b""
        return np.frombuffer(head + bytes([r % 251]) * (3 * n_blocks + r % 5) + tail, dtype=np.uint8)


def _shard_worker(rank, world, port, total_frames, tmp):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mrcaudiocodec_b200 import dist as mdist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    codec = _FakeShardCodec()
    nblk = (total_frames + codec.L - 1) // codec.L
    lo, hi = mdist.stream_shard_range(nblk, rank, world)
    f0, f1 = codec.shard_pcm_range(total_frames, lo, hi - lo)
    pcm = np.zeros((f1 - f0, 2), np.int16)
    blob, offsets = mdist.encode_stream_sharded(codec, pcm, f0, total_frames)
    np.save(os.path.join(tmp, "shard%d.npy" % rank), blob)
    np.save(os.path.join(tmp, "soff%d.npy" % rank), offsets)
    dist.destroy_process_group()


@pytest.mark.parametrize("total_frames", [10 * 1024 + 17, 1024, 3])
def test_stream_sharding_relay_world2(tmp_path, total_frames):
    """one stream over two ranks: block ranges partition the stream, the reservoir travels rank 0 -> rank 1 as one
    point-to-point message, the byte offsets come from the all-gather"""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_shard_worker, args=(2, port, total_frames, str(tmp_path)), nprocs=2, join=True)
    codec = _FakeShardCodec()
    nblk = (total_frames + 1023) // 1024
    box = [0]
    whole = codec.encode_shard(np.zeros((total_frames, 2), np.int16), 0, total_frames, 0, nblk, True, True, lambda: 0,
                               lambda r: box.__setitem__(0, r))
    parts = [np.load(os.path.join(str(tmp_path), "shard%d.npy" % r)) for r in range(2)]
    offs = [np.load(os.path.join(str(tmp_path), "soff%d.npy" % r)) for r in range(2)]
    assert np.array_equal(offs[0], offs[1]) and offs[0][0] == 0
    assert [int(offs[0][r + 1] - offs[0][r]) for r in range(2)] == [p.size for p in parts]
    # the second rank continued the first one's checksum: last byte value equals the single-shard run's
    assert parts[1][-1] == ord("T") and parts[0][0] == ord("H")
    assert int(parts[1][-2]) == int(whole[-2])
