"""Host-side shard arithmetic for multi-GPU runs.

The codec has no exchange step: decode shards by block range or by stream, encode by stream (and, in the
segment-parallel extension only, one stream by segment range), and "reassembly" is every rank writing
its disjoint slice.  These helpers only decide who owns what;
they are pure integer functions so the N>1 logic can be tested on CPU (tests/test_shard_gloo.py).
"""
from dataclasses import dataclass


def split_range(n_units, world, rank):
    """Contiguous, balanced split of range(n_units): the first (n_units % world) ranks get one more."""
    assert world >= 1 and 0 <= rank < world
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class BlockShard:
    """Blocks [block_begin, block_end) of one stream: what a rank needs to decode its part."""
    block_begin: int
    block_end: int
    byte_begin: int      # offset of the first block's first byte in the .aad file
    byte_end: int        # one past the last byte this rank reads (clipped to the file size)
    sample_begin: int    # first PCM sample (per channel) this rank produces
    sample_end: int


def decode_block_shard(num_samples, samples_per_block, block_size, file_size, world, rank, header_bytes=31):
    """Shard the decode of ONE stream by block range (src/aad_decoder.c:514-534: blocks are independent)."""
    n_blocks = (num_samples + samples_per_block - 1) // samples_per_block
    b0, b1 = split_range(n_blocks, world, rank)
    byte_begin = header_bytes + b0 * block_size
    byte_end = min(header_bytes + b1 * block_size, file_size)
    return BlockShard(b0, b1, byte_begin, max(byte_end, byte_begin),
                      min(b0 * samples_per_block, num_samples), min(b1 * samples_per_block, num_samples))


def encode_stream_shard(num_streams, world, rank):
    """Encode shards only across independent streams: one stream is a serial chain
    (src/aad_encoder.c:853-886 carries the predictor state from block to block)."""
    return split_range(num_streams, world, rank)


@dataclass(frozen=True)
class SegmentShard:
    """Whole segments [segment_begin, segment_end) of one stream for the segment-parallel encoder
    (AADGpuGroup_EncodeInterleaved16): the samples a rank reads, the blocks it
    encodes and the bytes of the stream it writes (the 31-byte file header travels with block 0)."""
    segment_begin: int
    segment_end: int
    block_begin: int
    block_end: int
    sample_begin: int
    sample_end: int
    byte_begin: int
    byte_end: int


def encode_segment_shard(num_samples, samples_per_block, block_size, stream_bytes, segment_blocks, world, rank,
                         header_bytes=31):
    """Shard the ENCODE of one stream by segment range.  Only with segment_blocks > 0: without segments a
    stream is one serial chain per channel (src/aad_encoder.c:853-886) and does not shard."""
    if segment_blocks <= 0:
        raise ValueError("one stream does not shard bit-exactly: segment_blocks must be > 0")
    n_blocks = (num_samples + samples_per_block - 1) // samples_per_block
    n_segments = (n_blocks + segment_blocks - 1) // segment_blocks
    g0, g1 = split_range(n_segments, world, rank)
    b0, b1 = min(g0 * segment_blocks, n_blocks), min(g1 * segment_blocks, n_blocks)
    byte_begin = (header_bytes + b0 * block_size) if b0 else 0
    byte_end = min(header_bytes + b1 * block_size, stream_bytes) if b1 > b0 else byte_begin
    return SegmentShard(g0, g1, b0, b1, min(b0 * samples_per_block, num_samples), min(b1 * samples_per_block, num_samples),
                        byte_begin, byte_end)
