#!/usr/bin/env python3
"""Regenerate tests/golden/golden.json from the COMPILED REFERENCE (oracle/_ref/*.so).

Run in the build container (where /root/reference exists):
    make -C oracle && python tests/golden/make_golden.py
Every entry is produced by calling the unmodified reference's AADEncoder_EncodeWhole /
AADDecoder_DecodeWhole through ctypes -- 1/2-channel cases on the stock build, 3..8-channel
cases on the build with src/aad.h:13 patched to 8.  The GPU box has no reference sources;
it checks the CUDA path against these hashes (and against the oracle / _ref libraries that
travel with the snapshot).

The .wav / .aad files in this directory are the reference's own test fixtures
(test/*.wav, test/sin300Hz*.aad, test/sin300Hz*_decoded.wav), copied verbatim.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import aadtest  # noqa: E402
from aad_b200.capi import OK, AADCApi  # noqa: E402


def cases():
    out = []
    for name in ("sin300Hz", "sin300Hz_mono", "bunny1", "pi_15-25sec", "unit_impulse", "unit_impulse_mono"):
        for bits in (2, 3, 4):
            for trials in (0, 2):
                out.append(dict(source=f"wav:{name}", bits=bits, max_block=1024, ms=False, trials=trials))
    # block-size / MS sweep of the reference's own wav test list (test/test_aad_encode_decode.c:508-616)
    for name in ("sin300Hz", "pi_15-25sec", "unit_impulse"):
        for block in (128, 256, 4096):
            for ms in (False, True):
                out.append(dict(source=f"wav:{name}", bits=4, max_block=block, ms=ms, trials=1))
        out.append(dict(source=f"wav:{name}", bits=4, max_block=1024, ms=True, trials=2))
        out.append(dict(source=f"wav:{name}", bits=3, max_block=1024, ms=True, trials=2))
        out.append(dict(source=f"wav:{name}", bits=2, max_block=256, ms=True, trials=0))
    # synthetic grid of test/test_aad_encode_decode.c:303-420 (+ ragged lengths, extra signals)
    for kind in aadtest.SIGNALS:
        for channels in (1, 2):
            for bits in (2, 3, 4):
                for block in (128, 1024):
                    for trials in (0, 1):
                        for ms in ((False, True) if channels == 2 else (False,)):
                            out.append(dict(source=f"signal:{kind}", channels=channels, n=2048, seed=1, rate=8000,
                                            bits=bits, max_block=block, ms=ms, trials=trials))
    for n in (1, 3, 4, 5, 6, 12, 13, 255, 2016, 2017, 2020, 2021, 4037, 5000):
        for bits in (2, 3, 4):
            out.append(dict(source="signal:music", channels=2, n=n, seed=n, rate=44100, bits=bits, max_block=1024,
                            ms=(n % 2 == 0), trials=2))
            out.append(dict(source="signal:noise", channels=1, n=n, seed=n, rate=44100, bits=bits, max_block=64,
                            ms=False, trials=1))
    # multichannel (patched reference)
    for channels in (3, 5, 8):
        for bits in (2, 3, 4):
            for kind in ("music", "noise"):
                out.append(dict(source=f"signal:{kind}", channels=channels, n=3000, seed=channels, rate=96000,
                                bits=bits, max_block=1024, ms=(channels == 8 and bits == 3), trials=2 if kind == "music" else 0))
    return out


def source_pcm(case):
    kind, name = case["source"].split(":")
    if kind == "wav":
        pcm, rate = aadtest.read_wav16(aadtest.GOLDEN / f"{name}.wav")
        return pcm, rate
    return aadtest.signal(name, case["channels"], case["n"], case["seed"]), case["rate"]


def main():
    ref = AADCApi(ROOT / "oracle" / "_ref" / "libaad_ref.so")
    ref8 = AADCApi(ROOT / "oracle" / "_ref" / "libaad_ref8.so")
    table = []
    for case in cases():
        pcm, rate = source_pcm(case)
        lib = ref if pcm.shape[0] <= 2 else ref8
        rc, data = lib.encode_whole(pcm, rate, case["bits"], case["max_block"], case["ms"], case["trials"])
        assert rc == OK, (case, rc)
        rc, dec, _ = lib.decode_whole(data)
        assert rc == OK, (case, rc)
        entry = dict(case)
        entry.update(channels=int(pcm.shape[0]), n=int(pcm.shape[1]), rate=int(rate), aad_size=len(data),
                     aad_sha=aadtest.sha(data), pcm_sha=aadtest.pcm_sha(dec.astype(np.int16)))
        table.append(entry)
    (aadtest.GOLDEN / "golden.json").write_text(json.dumps(table, indent=0) + "\n")
    print(f"wrote {len(table)} cases")


if __name__ == "__main__":
    main()
