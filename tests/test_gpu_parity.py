"""Parity of the CUDA path (through the C ABI of libaad_b200.so) with the reference -- GPU only.

Bar: bit-exact.  Identical .aad bytes on encode, identical PCM on decode, for every bit depth,
channel count, block size, MS setting and trial count the reference supports, on the reference's
own fixtures, on the golden table generated from the compiled reference, and differentially
against the oracle (and the compiled reference, when oracle/_ref travelled) on seeded inputs
including the pathological signals that wrap int32 (test/test_aad_encode_decode.c:447-470).
"""
import ctypes as C

import numpy as np
import pytest

import aadtest
from aad_b200 import capi
from aad_b200.capi import OK, make_param

pytestmark = pytest.mark.gpu


def _case_pcm(case):
    kind, name = case["source"].split(":")
    if kind == "wav":
        return aadtest.read_wav16(aadtest.GOLDEN / f"{name}.wav")
    return aadtest.signal(name, case["channels"], case["n"], case["seed"]), case["rate"]


# ---- the reference's own fixtures through the drop-in API ---------------------------------------

@pytest.mark.parametrize("stem", ["sin300Hz", "sin300Hz_mono"])
def test_shipped_fixtures_bit_exact(product, gpu_ctx, stem):
    api, _ = product
    pcm, rate = aadtest.read_wav16(aadtest.GOLDEN / f"{stem}.wav")
    golden_aad = (aadtest.GOLDEN / f"{stem}.aad").read_bytes()
    golden_dec, _ = aadtest.read_wav16(aadtest.GOLDEN / f"{stem}_decoded.wav")
    rc, data = api.encode_whole(pcm, rate, 4, 1024, False, 2)       # CLI defaults, test/make_test_data.sh:4-5
    assert rc == OK and data == golden_aad
    rc, dec, h = api.decode_whole(golden_aad)                        # test/test_aad_decoder.c:256-328
    assert rc == OK and h.num_channels == pcm.shape[0]
    assert np.array_equal(dec, golden_dec.astype(np.int32))


def test_golden_table_through_dropin_api(product, gpu_ctx):
    api, _ = product
    for case in aadtest.golden_table():
        pcm, rate = _case_pcm(case)
        rc, data = api.encode_whole(pcm, rate, case["bits"], case["max_block"], case["ms"], case["trials"])
        assert rc == OK, case
        assert len(data) == case["aad_size"] and aadtest.sha(data) == case["aad_sha"], case
        rc, dec, _ = api.decode_whole(data)
        assert rc == OK and aadtest.pcm_sha(dec.astype(np.int16)) == case["pcm_sha"], case


@pytest.mark.parametrize("generic_only", [0, 1])
def test_golden_table_through_batch_api(product, gpu_ctx, generic_only):
    """int16 batch entry points (the fast kernels when generic_only == 0), one stream per call and
    all same-shaped cases of a source in one call."""
    _, gpu = product
    gpu.lib.AADGpu_SetKernelPath(generic_only)
    try:
        for case in aadtest.golden_table():
            pcm, rate = _case_pcm(case)
            aad, sizes = gpu.encode_batch(gpu_ctx, pcm[None], rate, case["bits"], case["max_block"], case["ms"], case["trials"])
            data = aad[0, :sizes[0]].tobytes()
            assert len(data) == case["aad_size"] and aadtest.sha(data) == case["aad_sha"], case
            dec = gpu.decode_batch(gpu_ctx, aad, pcm.shape[1], rate, pcm.shape[0], case["bits"], case["max_block"],
                                   case["ms"], sizes=sizes)
            assert aadtest.pcm_sha(dec[0]) == case["pcm_sha"], case
    finally:
        gpu.lib.AADGpu_SetKernelPath(0)


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels", [1, 2])
def test_fast_and_generic_kernels_agree(product, gpu_ctx, bits, channels):
    """Odd block sizes, odd lengths, many blocks per stream: fast path vs generic path, byte for byte."""
    _, gpu = product
    rng = np.random.default_rng(bits + 10 * channels)
    for block in (64 * channels, 256, 1024, 4096):
        n_streams, n_max = 70, 9000
        lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
        lens[0] = n_max
        pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
        for i in range(n_streams):
            pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), 50 + i)
        res = []
        for generic_only in (0, 1):
            gpu.lib.AADGpu_SetKernelPath(generic_only)
            try:
                aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 48000, bits, block, channels == 2, 1, num_samples=lens)
                dec = gpu.decode_batch(gpu_ctx, aad, n_max, 48000, channels, bits, block, channels == 2, sizes=sizes)
            finally:
                gpu.lib.AADGpu_SetKernelPath(0)
            for i in range(n_streams):            # bytes past each stream's end are not part of the result
                aad[i, sizes[i]:] = 0
                dec[i, :, lens[i]:] = 0
            res.append((aad, sizes, dec))
        assert np.array_equal(res[0][1], res[1][1]), (bits, channels, block)
        assert np.array_equal(res[0][0], res[1][0]), (bits, channels, block)
        assert np.array_equal(res[0][2], res[1][2]), (bits, channels, block)


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels,ms", [(1, False), (2, False), (2, True)])
def test_decoder_tasks_spanning_streams(product, gpu_ctx, oracle, bits, channels, ms):
    """aad_decode_fast's warp tasks may run on from one stream into the next (kernel path 5: always, 6: never, 0: where
    per-stream tasks would idle lanes).  Ragged batches of short streams -- 1 .. 70 blocks, so a task holds pieces of up
    to 32 streams, partial last blocks in the middle of a task, empty and 1-sample streams, truncated data, an output
    buffer shorter than the stream -- decode to the same samples on every path, and those are the oracle's."""
    _, gpu = product
    rng = np.random.default_rng(4000 + bits * 10 + channels + int(ms))
    for block, n_max in ((64 * channels, 900), (256, 9000), (1024, 30000)):
        n_streams = 83
        lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
        lens[0], lens[1], lens[2], lens[5] = n_max, 1, 4, 5
        lens[10:20] = rng.integers(1, 60, size=10)          # many streams inside one task
        pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
        for i in range(n_streams):
            pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), 90 + i)
        aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, block, ms, 0, num_samples=lens)
        cut = sizes.copy()
        cut[3] = max(31, sizes[3] // 2)                      # data ends inside a block
        cut[4] = 31                                          # nothing but the file header
        res = {}
        for path in (5, 6, 0):
            gpu.lib.AADGpu_SetKernelPath(path)
            try:
                res[path] = (gpu.decode_batch(gpu_ctx, aad, n_max, 44100, channels, bits, block, ms, sizes=sizes),
                             gpu.decode_batch(gpu_ctx, aad, n_max, 44100, channels, bits, block, ms, sizes=cut))
            finally:
                gpu.lib.AADGpu_SetKernelPath(0)
        for i in range(n_streams):
            _, want, _ = oracle.decode(aad[i, :sizes[i]].tobytes())
            for path in (5, 6, 0):
                assert np.array_equal(res[path][0][i, :, :lens[i]], want), (block, path, i, lens[i])
        for path in (5, 0):
            assert np.array_equal(res[path][1], res[6][1]), (block, path)


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels", [1, 2, 3, 4, 5, 6, 7, 8])
def test_wide_decoder_agrees_with_generic_and_oracle(product, gpu_ctx, oracle, bits, channels):
    """aad_decode_wide (any channel count, staged through shared memory; mono / stereo forced onto it with
    kernel path 2) against the generic decoder on ragged batches and several block sizes, MS on and off,
    truncated streams; a sample of streams also against the oracle."""
    _, gpu = product
    rng = np.random.default_rng(300 + bits + 10 * channels)
    for block in (32 * channels, 128 * channels, 1024, 4096):
        for ms in ((False, True) if channels >= 2 else (False,)):
            n_streams, n_max = 45, 7000
            _, bs, spb = oracle.geometry(block, channels, bits)
            lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
            lens[0], lens[1], lens[2] = n_max, 4, 5
            pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
            for i in range(n_streams):
                pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), 90 + i)
            aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 48000, bits, block, ms, 1, num_samples=lens)
            cut = sizes.copy()                      # truncate a few streams inside / at the edge of a block
            for i in range(3, n_streams, 7):
                cut[i] = max(31, int(sizes[i]) - int(rng.integers(0, 2 * block)))
            res = []
            for path in (2, 1):
                gpu.lib.AADGpu_SetKernelPath(path)
                try:
                    full = gpu.decode_batch(gpu_ctx, aad, n_max, 48000, channels, bits, block, ms, sizes=sizes)
                    part = gpu.decode_batch(gpu_ctx, aad, n_max, 48000, channels, bits, block, ms, sizes=cut)
                finally:
                    gpu.lib.AADGpu_SetKernelPath(0)
                for i in range(n_streams):
                    full[i, :, lens[i]:] = 0
                    # a block is decoded when all of its channel headers are present (src/aad_decoder.c:347);
                    # what lies behind the last such block is not part of the result
                    nb = (int(cut[i]) - 31 - 18 * channels) // bs + 1 if cut[i] >= 31 + 18 * channels else 0
                    part[i, :, min(int(lens[i]), nb * spb):] = 0
                res.append((full, part))
            assert np.array_equal(res[0][0], res[1][0]), (bits, channels, block, ms)
            assert np.array_equal(res[0][1], res[1][1]), (bits, channels, block, ms)
            for i in range(0, n_streams, 9):
                _, want, _ = oracle.decode(aad[i, :sizes[i]].tobytes())
                assert np.array_equal(res[0][0][i, :, :lens[i]], want), (bits, channels, block, ms, i)


def test_five_bits_is_rejected_like_the_reference(product, gpu_ctx):
    api, gpu = product
    pcm = aadtest.signal("sine", 1, 500)
    assert api.encode_whole(pcm, 8000, 5)[0] == capi.INVALID_FORMAT          # src/aad_encoder.c:743-746
    assert api.encode_whole(pcm, 8000, 1)[0] == capi.INVALID_FORMAT          # src/aad_encoder.c:165-168
    with pytest.raises(Exception):
        gpu.encode_batch(gpu_ctx, pcm[None], 8000, 5)


# ---- differential against the oracle ------------------------------------------------------------

def _random_cases(n_cases, seed, max_channels=8):
    rng = np.random.default_rng(seed)
    for i in range(n_cases):
        ch = int(rng.choice([1, 1, 2, 2, 2, 3, 4, 8])) if max_channels > 2 else int(rng.integers(1, 3))
        yield dict(kind=aadtest.SIGNALS[int(rng.integers(len(aadtest.SIGNALS)))], ch=ch,
                   n=int(rng.choice([1, 3, 4, 5, 9, 100, 2016, 2017, 4100, 9001])) if i % 3 else int(rng.integers(1, 12000)),
                   bits=int(rng.integers(2, 5)), block=int(rng.choice([18 * ch + 8, 128, 256, 1024, 4096])) if ch <= 2 else 1024,
                   ms=bool(rng.integers(2)) and ch >= 2, trials=int(rng.integers(0, 4)), seed=1000 + i)


def test_dropin_api_matches_oracle_on_random_cases(product, gpu_ctx, oracle):
    api, _ = product
    for c in _random_cases(150, seed=21):
        pcm = aadtest.signal(c["kind"], c["ch"], c["n"], c["seed"])
        rc_o, want = oracle.encode(pcm, 44100, c["bits"], c["block"], c["ms"], c["trials"])
        rc_g, got = api.encode_whole(pcm, 44100, c["bits"], c["block"], c["ms"], c["trials"])
        assert rc_g == rc_o, c
        if rc_o != 0:
            continue
        assert got == want, c
        rc_o, dec_o, _ = oracle.decode(want)
        rc_g, dec_g, _ = api.decode_whole(want)
        assert rc_g == rc_o == 0 and np.array_equal(dec_g, dec_o.astype(np.int32)), c


def test_dropin_api_matches_compiled_reference(product, gpu_ctx, ref):
    api, _ = product
    for c in _random_cases(60, seed=22, max_channels=2):
        pcm = aadtest.signal(c["kind"], c["ch"], c["n"], c["seed"])
        rc_r, want = ref.encode_whole(pcm, 48000, c["bits"], c["block"], c["ms"], c["trials"])
        rc_g, got = api.encode_whole(pcm, 48000, c["bits"], c["block"], c["ms"], c["trials"])
        assert rc_g == rc_r and got == want, c
        if rc_r == OK:
            _, dec_r, _ = ref.decode_whole(want)
            _, dec_g, _ = api.decode_whole(want)
            assert np.array_equal(dec_g, dec_r), c


def test_dropin_and_batch_api_match_patched_reference_up_to_8_channels(product, gpu_ctx, ref8):
    """3, 4 and 8 channels directly against the compiled reference (the reference's sources with AAD_MAX_NUM_CHANNELS at 8,
    oracle/Makefile) instead of the oracle port: the drop-in calls, and the batch path's kernels on the same streams."""
    api, gpu = product
    cases = [c for c in _random_cases(120, seed=23, max_channels=8) if c["ch"] >= 3][:40]
    assert len(cases) >= 20 and {c["ch"] for c in cases} == {3, 4, 8}
    for c in cases:
        pcm = aadtest.signal(c["kind"], c["ch"], c["n"], c["seed"])
        rc_r, want = ref8.encode_whole(pcm, 48000, c["bits"], c["block"], c["ms"], c["trials"])
        rc_g, got = api.encode_whole(pcm, 48000, c["bits"], c["block"], c["ms"], c["trials"])
        assert rc_g == rc_r and got == want, c
        if rc_r != OK:
            continue
        _, dec_r, _ = ref8.decode_whole(want)
        _, dec_g, _ = api.decode_whole(want)
        assert np.array_equal(dec_g, dec_r), c
        # the batch path (fast / wide kernels) on the same stream
        aad, sizes = gpu.encode_batch(gpu_ctx, pcm[None].astype(np.int16), 48000, c["bits"], c["block"], c["ms"], c["trials"])
        assert aad[0, :sizes[0]].tobytes() == want, c
        dec_b = gpu.decode_batch(gpu_ctx, aad, c["n"], 48000, c["ch"], c["bits"], c["block"], c["ms"], sizes=sizes)
        assert np.array_equal(dec_b[0], dec_r), c


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels,ms", [(1, False), (2, False), (2, True), (8, False)])
def test_batch_api_matches_oracle(product, gpu_ctx, oracle, bits, channels, ms):
    """Many streams per call, ragged lengths, trials 2 (the CLI default)."""
    _, gpu = product
    n_streams, n_max = 37, 5200
    rng = np.random.default_rng(bits * 10 + channels)
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0], lens[1], lens[2] = n_max, 3, 4
    pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), i)
    aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, 1024, ms, 2, num_samples=lens)
    for i in range(n_streams):
        rc, want = oracle.encode(pcm[i, :, :lens[i]], 44100, bits, 1024, ms, 2)
        assert rc == 0 and sizes[i] == len(want), (i, lens[i])
        assert aad[i, :sizes[i]].tobytes() == want, (i, lens[i])
    dec = gpu.decode_batch(gpu_ctx, aad, n_max, 44100, channels, bits, 1024, ms, sizes=sizes)
    for i in range(n_streams):
        _, want, _ = oracle.decode(aad[i, :sizes[i]].tobytes())
        assert np.array_equal(dec[i, :, :lens[i]], want), (i, lens[i])


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels,ms", [(1, False), (2, True), (8, False)])
def test_encoder_pass_pairing_does_not_change_a_byte(product, gpu_ctx, oracle, bits, channels, ms):
    """With few chains the encoder interleaves the two independent dry passes of a block in one thread
    (enc_run_pair); with pairing off they run one after the other.  Ragged lengths (pairs of unequal
    length in every last block), trials 1..3, two block sizes; the paired result also against the oracle."""
    _, gpu = product
    n_streams, n_max = 41, 4300
    rng = np.random.default_rng(1000 + bits * 10 + channels)
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0], lens[1], lens[2], lens[3] = n_max, 3, 4, 5
    pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), 7 * i)
    for block in (256 * channels // (2 if channels == 8 else 1), 1024):
        for trials in (1, 2, 3):
            res = []
            for pairing in (1, 0):
                gpu.lib.AADGpu_SetEncoderPairing(pairing)
                try:
                    aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, block, ms, trials, num_samples=lens)
                finally:
                    gpu.lib.AADGpu_SetEncoderPairing(1)
                for i in range(n_streams):
                    aad[i, sizes[i]:] = 0
                res.append((aad, sizes))
            assert np.array_equal(res[0][1], res[1][1]), (block, trials)
            assert np.array_equal(res[0][0], res[1][0]), (block, trials)
            for i in range(0, n_streams, 5):
                rc, want = oracle.encode(pcm[i, :, :lens[i]], 44100, bits, block, ms, trials)
                assert rc == 0 and res[0][0][i, :res[0][1][i]].tobytes() == want, (block, trials, i)


@pytest.mark.parametrize("bits", [2, 3, 4])
@pytest.mark.parametrize("channels,ms", [(1, False), (2, True), (2, False), (8, False)])
def test_encoder_schedules_do_not_change_a_byte(product, gpu_ctx, oracle, bits, channels, ms):
    """The encoder's pass schedules for scarce chains (aad_encode_roles.cuh): 0 = one pass at a time, 2 = two dry
    passes interleaved in one thread, 3 = helper lanes run the baseline passes, 4 = helper lanes + a second warp
    emitting ahead of the decision, 1 = chosen by shape.  Same bytes from all of them, and the oracle's: ragged
    batches (chains end at different blocks), one and three streams, trials 1..3 (3 falls back to the plain passes)."""
    _, gpu = product
    rng = np.random.default_rng(2000 + bits * 10 + channels)
    for n_streams, n_max in ((37, 4300), (1, 9000), (3, 2500)):
        lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
        lens[0] = n_max
        if n_streams > 4:
            lens[1], lens[2], lens[3] = 3, 4, 5
        pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
        for i in range(n_streams):
            pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[(i + bits) % len(aadtest.SIGNALS)], channels, int(lens[i]), 11 * i)
        for block in (256 * channels // (2 if channels == 8 else 1), 1024):
            for trials in (1, 2, 3):
                res = {}
                for schedule in (0, 1, 2, 3, 4):
                    gpu.lib.AADGpu_SetEncoderSchedule(schedule)
                    try:
                        aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, block, ms, trials, num_samples=lens)
                    finally:
                        gpu.lib.AADGpu_SetEncoderSchedule(1)
                    for i in range(n_streams):
                        aad[i, sizes[i]:] = 0
                    res[schedule] = (aad, sizes)
                for schedule in (1, 2, 3, 4):
                    assert np.array_equal(res[0][1], res[schedule][1]), (n_streams, block, trials, schedule)
                    assert np.array_equal(res[0][0], res[schedule][0]), (n_streams, block, trials, schedule)
                for i in range(0, n_streams, 6):
                    rc, want = oracle.encode(pcm[i, :, :lens[i]], 44100, bits, block, ms, trials)
                    assert rc == 0 and res[0][0][i, :res[0][1][i]].tobytes() == want, (n_streams, block, trials, i)


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_quantiser_near_silence_matches_oracle(product, gpu_ctx, oracle, bits):
    """The encoder's shift-free quantiser has two copies of the 16-sample loop: one for units whose step index stays
    clear of the first rows of the step table (steps <= 2^(b-2)), one for units that may reach them (EncQuant in
    aad_encode_fast.cuh).  Signals that fade into silence and back, +-1..6 LSB noise, silence and an impulse walk
    the index across that boundary in both directions, in every schedule, mono and stereo."""
    _, gpu = product
    for channels, ms in ((1, False), (2, True)):
        n_streams, n_max = 45, 6100
        rng = np.random.default_rng(3000 + bits * 10 + channels)
        lens = rng.integers(200, n_max + 1, size=n_streams).astype(np.uint32)
        lens[0] = n_max
        pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
        for i in range(n_streams):
            pcm[i, :, :lens[i]] = aadtest.signal(aadtest.QUIET_SIGNALS[i % 2 if i < 40 else 2 + i % 2], channels, int(lens[i]), i)
        for schedule in (1, 0, 2, 3):
            gpu.lib.AADGpu_SetEncoderSchedule(schedule)
            try:
                aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 44100, bits, 1024, ms, 2, num_samples=lens)
            finally:
                gpu.lib.AADGpu_SetEncoderSchedule(1)
            for i in range(n_streams):
                rc, want = oracle.encode(pcm[i, :, :lens[i]], 44100, bits, 1024, ms, 2)
                assert rc == 0 and aad[i, :sizes[i]].tobytes() == want, (channels, schedule, i, lens[i])


def test_encoder_schedules_carry_state_across_launch_slices(product, gpu_ctx, oracle):
    """AADGpu_EncodeBatch cuts a batch larger than 64 MiB of PCM into block-range slices; the chain state crosses the
    launches through the device state array in every schedule (first-block rules only in the stream's block 0)."""
    _, gpu = product
    n_streams, channels, n = 40, 2, 450_000           # 72 MB of PCM: two slices
    pcm = np.stack([aadtest.signal("music", channels, n, i) for i in range(n_streams)])
    res = {}
    for schedule in (0, 3, 4):
        gpu.lib.AADGpu_SetEncoderSchedule(schedule)
        try:
            res[schedule] = gpu.encode_batch(gpu_ctx, pcm, 44100, 4, 1024, False, 2)
        finally:
            gpu.lib.AADGpu_SetEncoderSchedule(1)
    for schedule in (3, 4):
        assert np.array_equal(res[0][0], res[schedule][0]) and np.array_equal(res[0][1], res[schedule][1]), schedule
    rc, want = oracle.encode(pcm[7], 44100, 4, 1024, False, 2)
    assert rc == 0 and res[0][0][7, :res[0][1][7]].tobytes() == want


@pytest.mark.parametrize("bits,channels,ms", [(4, 1, False), (3, 2, True), (2, 8, False)])
def test_reconstruct_batch_equals_encode_then_decode(product, gpu_ctx, bits, channels, ms):
    """AADGpu_ReconstructBatch = AADGpu_EncodeBatch + AADGpu_DecodeBatch in one sliced pass (ragged lengths,
    enough samples for several slices); with and without asking for the .aad."""
    _, gpu = product
    n_streams, n_max = 700, 30000
    rng = np.random.default_rng(bits)
    lens = rng.integers(1, n_max + 1, size=n_streams).astype(np.uint32)
    lens[0] = n_max
    pcm = np.zeros((n_streams, channels, n_max), dtype=np.int16)
    for i in range(n_streams):
        pcm[i, :, :lens[i]] = aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], channels, int(lens[i]), i)
    want_aad, want_sizes = gpu.encode_batch(gpu_ctx, pcm, 32000, bits, 1024, ms, 2, num_samples=lens)
    want_pcm = gpu.decode_batch(gpu_ctx, want_aad, n_max, 32000, channels, bits, 1024, ms, sizes=want_sizes)
    b = gpu.batch(n_streams, n_max, make_param(channels, 32000, bits, 1024, ms, 2))
    for with_aad in (True, False):
        aad = np.zeros_like(want_aad)
        sizes = np.zeros(n_streams, dtype=np.uint32)
        out = np.zeros_like(want_pcm)
        rc = gpu.lib.AADGpu_ReconstructBatch(gpu_ctx, C.byref(b), pcm.ctypes.data, lens.ctypes.data,
                                             aad.ctypes.data if with_aad else None, sizes.ctypes.data if with_aad else None,
                                             out.ctypes.data)
        assert rc == OK, gpu.last_error()
        for i in range(n_streams):
            assert np.array_equal(out[i, :, :lens[i]], want_pcm[i, :, :lens[i]]), i
            if with_aad:
                assert sizes[i] == want_sizes[i] and np.array_equal(aad[i, :sizes[i]], want_aad[i, :sizes[i]]), i


def test_trials_zero_and_many(product, gpu_ctx, oracle):
    _, gpu = product
    pcm = np.stack([aadtest.signal(k, 2, 7000, 5) for k in aadtest.SIGNALS])
    for trials in (0, 1, 5):
        aad, sizes = gpu.encode_batch(gpu_ctx, pcm, 48000, 4, 1024, False, trials)
        for i in range(len(pcm)):
            _, want = oracle.encode(pcm[i], 48000, 4, 1024, False, trials)
            assert aad[i, :sizes[i]].tobytes() == want, (trials, i)


# ---- API semantics that need the device -----------------------------------------------------------

def test_handle_reuse_carries_state(product, gpu_ctx, oracle):
    """src/aad_encoder.c:299-301,797-799: weights survive between EncodeWhole calls on one handle."""
    api, _ = product
    a, b = aadtest.signal("music", 2, 3000, 1), aadtest.signal("sine", 2, 2500, 2)
    h = api.lib.AADEncoder_Create(1024, None, 0)
    assert api.lib.AADEncoder_SetEncodeParameter(h, C.byref(make_param(2, 44100, 4, 1024, False, 1))) == OK
    _, ga = api.encode_whole(a, 44100, 4, handle=h)
    _, gb = api.encode_whole(b, 44100, 4, handle=h)
    # SetEncodeParameter resets the step index but not the weights
    assert api.lib.AADEncoder_SetEncodeParameter(h, C.byref(make_param(2, 44100, 4, 1024, False, 1))) == OK
    _, gc = api.encode_whole(a, 44100, 4, handle=h)
    api.lib.AADEncoder_Destroy(h)
    state = [[0] * 5, [0] * 5]
    _, oa = oracle.encode(a, 44100, 4, 1024, False, 1, state=state)
    _, ob = oracle.encode(b, 44100, 4, 1024, False, 1, state=state)
    for s in state:
        s[4] = 0
    _, oc = oracle.encode(a, 44100, 4, 1024, False, 1, state=state)
    assert ga == oa and gb == ob and gc == oc


def test_dropin_handles_are_independent_across_threads(product, gpu_ctx, oracle):
    """The reference has no shared mutable state: distinct handles may be used from distinct threads
    at the same time.  Here every handle of the process shares the default device context, whose entry points
    serialise on a per-context lock: 8 threads x 12 encode + decode calls of different shapes, every result against
    the oracle.  (ctypes releases the GIL during the calls, so they do overlap.)"""
    from concurrent.futures import ThreadPoolExecutor
    api, gpu = product
    rng = np.random.default_rng(77)
    cases = []
    for i in range(96):
        ch = int(rng.choice([1, 2, 2, 8]))
        cases.append(dict(pcm=aadtest.signal(aadtest.SIGNALS[i % len(aadtest.SIGNALS)], ch, int(rng.integers(1, 9000)), 300 + i),
                          bits=int(rng.integers(2, 5)), ms=bool(rng.integers(2)) and ch >= 2, trials=int(rng.integers(0, 3))))
    want = []
    for c in cases:
        rc, data = oracle.encode(c["pcm"], 44100, c["bits"], 1024, c["ms"], c["trials"])
        assert rc == 0
        want.append((data, oracle.decode(data)[1]))

    def work(i):
        c = cases[i]
        rc, data = api.encode_whole(c["pcm"], 44100, c["bits"], 1024, c["ms"], c["trials"])
        rc2, dec, _ = api.decode_whole(want[i][0])
        return rc, data, rc2, dec

    with ThreadPoolExecutor(8) as pool:
        got = list(pool.map(work, range(len(cases))))
    for i, (rc, data, rc2, dec) in enumerate(got):
        assert rc == 0 and data == want[i][0], i
        assert rc2 == 0 and np.array_equal(dec, want[i][1].astype(np.int32)), i
    # explicit contexts too: one shared by all threads
    pcm = np.stack([aadtest.signal("music", 1, 5000, s) for s in range(6)])

    def batch(i):
        aad, sizes = gpu.encode_batch(gpu_ctx, pcm[i:i + 3], 44100, 4, 1024, False, 1)
        return [aad[k, :sizes[k]].tobytes() for k in range(3)]

    with ThreadPoolExecutor(4) as pool:
        res = list(pool.map(batch, [0, 1, 2, 3] * 4))
    for j, r in enumerate(res):
        i = [0, 1, 2, 3][j % 4]
        for k in range(3):
            assert r[k] == oracle.encode(pcm[i + k], 44100, 4, 1024, False, 1)[1], (j, k)


def test_decode_block_and_oversized_buffer(product, gpu_ctx, oracle):
    api, _ = product
    pcm = aadtest.signal("music", 2, 3000, 7)
    _, data = oracle.encode(pcm, 44100, 3, 1024, True, 1)
    rc, h = api.decode_header(data)
    dec = api.lib.AADDecoder_Create(None, 0)
    assert api.lib.AADDecoder_SetHeader(dec, C.byref(h)) == OK
    _, want, _ = oracle.decode(data)
    buf = np.frombuffer(data, dtype=np.uint8)
    spb, bs = h.num_samples_per_block, h.block_size
    for b, room in [(0, spb), (1, spb), (1, 100), (2, 3000 - 2 * spb), (0, 3)]:
        out = np.full((2, spb), -7, dtype=np.int32)
        n = C.c_uint32()
        blk = buf[31 + b * bs:31 + (b + 1) * bs]
        rc = api.lib.AADDecoder_DecodeBlock(dec, blk.ctypes.data_as(C.POINTER(C.c_uint8)), len(blk),
                                            capi._planar_pointers([out[0], out[1]]), 2, room, C.byref(n))
        assert rc == OK and n.value == min(spb, room)
        assert np.array_equal(out[:, :n.value], want[:, b * spb:b * spb + n.value].astype(np.int32))
        assert (out[:, n.value:] == -7).all()
    short = buf[31:31 + 20]
    n = C.c_uint32()
    out = np.zeros((2, spb), dtype=np.int32)
    assert api.lib.AADDecoder_DecodeBlock(dec, short.ctypes.data_as(C.POINTER(C.c_uint8)), 20,
                                          capi._planar_pointers([out[0], out[1]]), 2, spb, C.byref(n)) == capi.INSUFFICIENT_DATA
    api.lib.AADDecoder_Destroy(dec)
    # DecodeWhole with a buffer larger than the stream: only num_samples are defined by the format;
    # the oracle and the product agree on the whole buffer when the tail reads as zero bytes
    rc_g, big_g, _ = api.decode_whole(data, buf_samples=3000 + 50, fill=-9)
    rc_o, big_o, _ = oracle.decode(data, buf_samples=3000 + 50, fill=-9)
    assert rc_g == rc_o == 0 and np.array_equal(big_g, big_o.astype(np.int32))


def test_truncated_stream(product, gpu_ctx, oracle):
    api, _ = product
    pcm = aadtest.signal("music", 1, 9000, 3)
    _, data = oracle.encode(pcm, 44100, 4, 1024, False, 0)
    _, full, _ = oracle.decode(data)
    # cut inside block 2's header: blocks 0,1 decode, then INSUFFICIENT_DATA (src/aad_decoder.c:347,522-527)
    cut = data[:31 + 2 * 1024 + 10]
    rc_g, dec_g, _ = api.decode_whole(cut, fill=-5)
    rc_o, dec_o, _ = oracle.decode(cut, fill=-5)
    assert rc_g == rc_o == capi.INSUFFICIENT_DATA
    assert np.array_equal(dec_g, dec_o.astype(np.int32))
    assert np.array_equal(dec_g[:, :2 * 2016], full[:, :2 * 2016].astype(np.int32))
    # cut at a block boundary: a clean shorter decode
    cut = data[:31 + 3 * 1024]
    rc_g, dec_g, _ = api.decode_whole(cut, fill=-5)
    rc_o, dec_o, _ = oracle.decode(cut, fill=-5)
    assert rc_g == rc_o == 0 and np.array_equal(dec_g, dec_o.astype(np.int32))


# ---- device-resident entry points (torch only provides the memory) ---------------------------------

def test_device_resident_roundtrip_and_synth(product, gpu_ctx, oracle):
    import torch
    from aad_b200.synth import synth_pcm16
    _, gpu = product
    n_streams, ch, n = 64, 2, 30000
    prm = make_param(ch, 44100, 4, 1024, False, 2)
    b = gpu.batch(n_streams, n, prm)
    dev = torch.device("cuda:0")
    pcm = torch.zeros((n_streams, ch, n), dtype=torch.int16, device=dev)
    aad = torch.zeros((n_streams, b.aad_stream_stride), dtype=torch.uint8, device=dev)
    sizes = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    out = torch.zeros_like(pcm)
    stream = torch.cuda.current_stream().cuda_stream
    launches = gpu.launch_count()
    assert gpu.lib.AADGpu_SynthBatchDevice(gpu_ctx, C.byref(b), 100, pcm.data_ptr(), stream) == OK
    assert gpu.lib.AADGpu_EncodeBatchDevice(gpu_ctx, C.byref(b), pcm.data_ptr(), None, aad.data_ptr(), sizes.data_ptr(),
                                            stream) == OK
    assert gpu.lib.AADGpu_DecodeBatchDevice(gpu_ctx, C.byref(b), aad.data_ptr(), None, out.data_ptr(), stream) == OK
    torch.cuda.synchronize()
    assert gpu.launch_count() - launches == 3
    host = pcm.cpu().numpy()
    assert np.array_equal(host, synth_pcm16(gpu.synth_lut(), 100, n_streams, ch, n, 44100))
    aad_h, sizes_h, out_h = aad.cpu().numpy(), sizes.cpu().numpy(), out.cpu().numpy()
    for i in (0, 1, 31, 63):
        _, want = oracle.encode(host[i], 44100, 4, 1024, False, 2)
        assert sizes_h[i] == len(want) and aad_h[i, :len(want)].tobytes() == want
        _, dec, _ = oracle.decode(want)
        assert np.array_equal(out_h[i], dec)


def test_interleave_helpers(product, gpu_ctx):
    import torch
    _, gpu = product
    x = torch.randint(-32768, 32767, (5000, 3), dtype=torch.int16, device="cuda:0")
    planar = torch.zeros((3, 5000), dtype=torch.int16, device="cuda:0")
    back = torch.zeros_like(x)
    s = torch.cuda.current_stream().cuda_stream
    assert gpu.lib.AADGpu_Deinterleave16Device(gpu_ctx, x.data_ptr(), planar.data_ptr(), 5000, 3, 5000, s) == OK
    assert gpu.lib.AADGpu_Interleave16Device(gpu_ctx, planar.data_ptr(), 5000, back.data_ptr(), 3, 5000, s) == OK
    torch.cuda.synchronize()
    assert torch.equal(planar, x.t().contiguous()) and torch.equal(back, x)
