"""The N>1 path on CPU: world_size-2 gloo processes shard a batch by stream (encode) and a long
stream by block range (decode), each rank works only on its slice, rank 0 reassembles, and the
result equals the single-process result.  The codec work is done by the oracle here (no GPU in
this container); what is under test is the sharding / reassembly logic of aad_b200/shard.py and
the barrier + max-over-ranks timing pattern bench.py uses."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import aadtest
    from aad_b200.shard import decode_block_shard, encode_stream_shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = aadtest.Oracle(ROOT / "oracle" / "liboracle.so")

    # ---- encode: streams sharded across ranks, sizes gathered, bytes reassembled in order ----
    n_streams, ch, n = 7, 2, 5000
    pcm = np.stack([aadtest.signal("music", ch, n, seed=i) for i in range(n_streams)])
    s0, s1 = encode_stream_shard(n_streams, world, rank)
    mine = [oracle.encode(pcm[i], 44100, 4, 1024, False, 1)[1] for i in range(s0, s1)]
    gathered = [None] * world
    dist.all_gather_object(gathered, (s0, mine))
    # ---- decode: ONE long stream sharded by block range ----------------------------------------
    long_pcm = aadtest.signal("music", 2, 40000, seed=99)
    _, data = oracle.encode(long_pcm, 48000, 3, 1024, True, 0)
    rc, whole, info = oracle.decode(data)
    sh = decode_block_shard(info.num_samples, info.samples_per_block, info.block_size, len(data), world, rank)
    # a rank only needs the 31-byte header and its own byte range: rebuild a private stream from them
    part_samples = sh.sample_end - sh.sample_begin
    hdr = bytearray(data[:31])
    hdr[14:18] = part_samples.to_bytes(4, "big")
    rc, part, _ = oracle.decode(bytes(hdr) + data[sh.byte_begin:sh.byte_end])
    assert rc == 0
    parts = [None] * world
    dist.all_gather_object(parts, (sh.sample_begin, part))
    # ---- timing pattern: barrier, local time, MAX over ranks ----------------------------------
    dist.barrier()
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        streams = [None] * n_streams
        for first, chunk in gathered:
            for k, blob in enumerate(chunk):
                streams[first + k] = blob
        want = [oracle.encode(pcm[i], 44100, 4, 1024, False, 1)[1] for i in range(n_streams)]
        out = np.zeros_like(whole)
        for begin, piece in parts:
            out[:, begin:begin + piece.shape[1]] = piece
        ok = streams == want and np.array_equal(out, whole) and float(t[0]) == float(world)
        (Path(out_dir) / "result.txt").write_text("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_sharding_reassembles_exactly(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "result.txt").read_text() == "ok"


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (100000, 8), (219, 8), (174194, 8)])
def test_split_range_is_a_partition(n, world):
    from aad_b200.shard import split_range
    pieces = [split_range(n, world, r) for r in range(world)]
    assert pieces[0][0] == 0 and pieces[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    sizes = [e - b for b, e in pieces]
    assert max(sizes) - min(sizes) <= 1


def test_block_shards_cover_the_file():
    from aad_b200.shard import decode_block_shard
    ns, spb, bs = 172_800_000, 992, 1024          # BASELINE config 3: 1 h, 48 kHz stereo 4-bit
    size = 31 + (ns // spb) * bs + 36 + ((ns % spb - 4 + 1) // 2) * 2
    shards = [decode_block_shard(ns, spb, bs, size, 8, r) for r in range(8)]
    assert shards[0].block_begin == 0 and shards[-1].block_end == 174_194
    assert shards[0].byte_begin == 31 and shards[-1].byte_end == size
    assert shards[-1].sample_end == ns
    for a, b in zip(shards, shards[1:]):
        assert a.block_end == b.block_begin and a.byte_end == b.byte_begin and a.sample_end == b.sample_begin


def _segment_worker(rank, world, port, out_dir):
    """One stream ENCODED by two ranks in the segment-parallel extension: every rank encodes its own segment range
    (the oracle on a fresh handle per segment = what the CUDA kernel's segment chains do), rank 0 reassembles."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import aadtest
    from aad_b200.shard import encode_segment_shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = aadtest.Oracle(ROOT / "oracle" / "liboracle.so")
    ch, n, bits, seg_blocks = 2, 61_003, 4, 3
    pcm = aadtest.signal("music", ch, n, seed=5)
    _, bs, spb = oracle.geometry(1024, ch, bits)
    _, whole = oracle.encode(pcm, 48000, bits, 1024, False, 0)          # only for the file header and the size
    sh = encode_segment_shard(n, spb, bs, len(whole), seg_blocks, world, rank)
    mine = bytearray()
    for g in range(sh.segment_begin, sh.segment_end):
        s0 = g * seg_blocks * spb
        rc, seg = oracle.encode(pcm[:, s0:s0 + seg_blocks * spb], 48000, bits, 1024, False, 2)
        assert rc == 0
        mine += seg[31:]
    if sh.block_begin == 0:
        mine = bytearray(whole[:31]) + mine
    assert len(mine) == sh.byte_end - sh.byte_begin
    parts = [None] * world
    dist.all_gather_object(parts, (sh.byte_begin, bytes(mine)))
    if rank == 0:
        out = bytearray(len(whole))
        for begin, blob in parts:
            out[begin:begin + len(blob)] = blob
        want = bytearray(whole[:31])
        for s0 in range(0, n, seg_blocks * spb):
            want += oracle.encode(pcm[:, s0:s0 + seg_blocks * spb], 48000, bits, 1024, False, 2)[1][31:]
        rc, dec, _ = oracle.decode(bytes(out))                           # a valid stream for the stock decoder
        ok = bytes(out) == bytes(want) and rc == 0 and dec.shape == pcm.shape
        (Path(out_dir) / "segments.txt").write_text("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_segment_encode_reassembles_exactly(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_segment_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "segments.txt").read_text() == "ok"


def test_segment_shards_cover_the_stream():
    from aad_b200.shard import encode_segment_shard
    ns, spb, bs = 172_800_000, 992, 1024          # BASELINE config 3
    size = 31 + (ns // spb) * bs + 36 + ((ns % spb - 4 + 1) // 2) * 2
    for world in (1, 2, 4, 8):
        shards = [encode_segment_shard(ns, spb, bs, size, 64, world, r) for r in range(world)]
        assert shards[0].byte_begin == 0 and shards[-1].byte_end == size and shards[-1].sample_end == ns
        assert shards[-1].block_end == 174_194 and shards[-1].segment_end == (174_194 + 63) // 64
        for a, b in zip(shards, shards[1:]):
            assert a.segment_end == b.segment_begin and a.block_end == b.block_begin and a.block_end % 64 == 0
            assert a.sample_end == b.sample_begin and a.byte_end == b.byte_begin
    # fewer segments than ranks: the extra ranks get empty shards
    few = [encode_segment_shard(3000, 992, 1024, 31 + 4 * 1024, 64, 4, r) for r in range(4)]
    assert few[0].block_end == 4 and all(s.block_begin == s.block_end for s in few[1:])
    with pytest.raises(ValueError):
        encode_segment_shard(ns, spb, bs, size, 0, 2, 0)
