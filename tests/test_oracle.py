"""Pins the oracle (oracle/aad_oracle.c) -- CPU only.

1. against the fixtures the reference ships and its own tests pin
   (test/test_aad_decoder.c:307-316,336-337; test/make_test_data.sh:4-7),
2. against tests/golden/golden.json (hashes produced by the compiled reference),
3. against the compiled reference itself (oracle/_ref) on fresh random cases, when present.
"""
import numpy as np
import pytest

import aadtest
from aad_b200.capi import OK


def _case_pcm(case):
    kind, name = case["source"].split(":")
    if kind == "wav":
        return aadtest.read_wav16(aadtest.GOLDEN / f"{name}.wav")
    return aadtest.signal(name, case["channels"], case["n"], case["seed"]), case["rate"]


@pytest.mark.parametrize("stem", ["sin300Hz", "sin300Hz_mono"])
def test_oracle_reproduces_shipped_fixtures(oracle, stem):
    pcm, rate = aadtest.read_wav16(aadtest.GOLDEN / f"{stem}.wav")
    golden_aad = (aadtest.GOLDEN / f"{stem}.aad").read_bytes()
    golden_dec, _ = aadtest.read_wav16(aadtest.GOLDEN / f"{stem}_decoded.wav")
    # encode with the CLI defaults (4-bit, block 1024, trials 2, no MS): identical bytes
    rc, data = oracle.encode(pcm, rate, 4, 1024, False, 2)
    assert rc == 0 and data == golden_aad
    # decode the shipped .aad: identical samples (the reference test compares sample << 16)
    rc, dec, info = oracle.decode(golden_aad)
    assert rc == 0 and info.channels == pcm.shape[0] and info.num_samples == 24000
    assert np.array_equal(dec, golden_dec)


def test_oracle_matches_golden_table(oracle):
    table = aadtest.golden_table()
    assert len(table) > 300
    for case in table:
        pcm, rate = _case_pcm(case)
        rc, data = oracle.encode(pcm, rate, case["bits"], case["max_block"], case["ms"], case["trials"])
        assert rc == 0, case
        assert len(data) == case["aad_size"] and aadtest.sha(data) == case["aad_sha"], case
        rc, dec, _ = oracle.decode(data)
        assert rc == 0 and aadtest.pcm_sha(dec) == case["pcm_sha"], case


# known answers from test/test_aad_encoder.c:33-57 plus SURVEY.md 8(a)
BLOCK_SIZE_KAT = [
    (32, 1, 4, 32, 32), (64, 2, 4, 64, 32), (64, 1, 3, 63, 124), (64, 2, 3, 60, 36), (128, 1, 3, 126, 292),
    (128, 2, 3, 126, 124), (1024, 1, 3, 1023, 2684), (1024, 2, 3, 1020, 1316), (32, 1, 2, 32, 60),
    (64, 1, 2, 64, 188), (64, 2, 2, 64, 60), (1024, 1, 4, 1024, 2016), (1024, 2, 4, 1024, 992),
    (1024, 1, 2, 1024, 4028), (1024, 2, 2, 1024, 1980), (1024, 8, 3, 1008, 292), (1024, 8, 4, 1024, 224),
    (1024, 8, 2, 1024, 444),
]


@pytest.mark.parametrize("max_block,ch,bits,bs,spb", BLOCK_SIZE_KAT)
def test_oracle_block_geometry(oracle, max_block, ch, bits, bs, spb):
    assert oracle.geometry(max_block, ch, bits) == (0, bs, spb)


def test_oracle_rejects_what_the_reference_rejects(oracle):
    assert oracle.geometry(17, 1, 4)[0] == 2          # block header does not fit
    assert oracle.geometry(32, 0, 4)[0] == 2
    assert oracle.geometry(32, 1, 0)[0] == 2
    assert oracle.geometry(32, 1, 5)[0] == 2          # 5 bits/sample: no such format (src/aad.h:19)
    pcm = aadtest.signal("sine", 1, 100)
    assert oracle.encode(pcm, 8000, 5)[0] == 2
    assert oracle.encode(pcm, 8000, 1)[0] == 2        # passes SetEncodeParameter, fails EncodeHeader
    assert oracle.encode(pcm, 8000, 4, ms=True)[0] == 2   # MS needs 2 channels
    assert oracle.decode(b"AAE\0" + bytes(40))[0] == 2
    assert oracle.decode(b"AAD\0" + bytes(10))[0] == 4    # shorter than the 31-byte header


def _random_cases(n_cases, max_channels, seed):
    rng = np.random.default_rng(seed)
    for i in range(n_cases):
        ch = int(rng.integers(1, max_channels + 1))
        yield dict(kind=aadtest.SIGNALS[int(rng.integers(len(aadtest.SIGNALS)))], ch=ch,
                   n=int(rng.integers(1, 7000)), bits=int(rng.integers(2, 5)),
                   block=int(rng.choice([18 * ch + 8, 64 * ch, 128, 256, 1024, 4096])) if ch <= 2 else 1024,
                   ms=bool(rng.integers(2)) and ch >= 2, trials=int(rng.integers(0, 4)), seed=i)


def _differential(oracle, lib, max_channels, seed, n_cases=120):
    for c in _random_cases(n_cases, max_channels, seed):
        pcm = aadtest.signal(c["kind"], c["ch"], c["n"], c["seed"])
        rc_r, data_r = lib.encode_whole(pcm, 44100, c["bits"], c["block"], c["ms"], c["trials"])
        rc_o, data_o = oracle.encode(pcm, 44100, c["bits"], c["block"], c["ms"], c["trials"])
        assert rc_r == rc_o, c
        if rc_r != OK:
            continue
        assert data_r == data_o, c
        rc_r, dec_r, _ = lib.decode_whole(data_r)
        rc_o, dec_o, _ = oracle.decode(data_r)
        assert rc_r == rc_o == 0 and np.array_equal(dec_r, dec_o), c


def test_oracle_vs_compiled_reference(oracle, ref):
    _differential(oracle, ref, 2, seed=11)


def test_oracle_vs_compiled_reference_wrapv(oracle, ref_wrapv):
    # the stock build relies on signed overflow wrapping; the -fwrapv build must agree with it
    _differential(oracle, ref_wrapv, 2, seed=11)


def test_oracle_vs_patched_reference_8ch(oracle, ref8):
    _differential(oracle, ref8, 8, seed=12, n_cases=60)


def test_oracle_handle_reuse_carries_state(oracle, ref):
    """Weights survive between EncodeWhole calls on one handle (src/aad_encoder.c:299-301,797-799)."""
    import ctypes as C
    from aad_b200.capi import make_param
    a = aadtest.signal("music", 2, 3000, 1)
    b = aadtest.signal("sine", 2, 2500, 2)
    h = ref.lib.AADEncoder_Create(1024, None, 0)
    assert ref.lib.AADEncoder_SetEncodeParameter(h, C.byref(make_param(2, 44100, 4, 1024, False, 1))) == OK
    _, ra = ref.encode_whole(a, 44100, 4, handle=h)
    _, rb = ref.encode_whole(b, 44100, 4, handle=h)
    ref.lib.AADEncoder_Destroy(h)
    state = [[0, 0, 0, 0, 0], [0, 0, 0, 0, 0]]
    _, oa = oracle.encode(a, 44100, 4, 1024, False, 1, state=state)
    _, ob = oracle.encode(b, 44100, 4, 1024, False, 1, state=state)
    assert ra == oa and rb == ob
    _, fresh = oracle.encode(b, 44100, 4, 1024, False, 1)
    assert fresh != ob      # the carried state really matters
