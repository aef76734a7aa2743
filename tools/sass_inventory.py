#!/usr/bin/env python3
"""Static SASS mnemonic counts per kernel of aad_b200/csrc/build/aad_kernels.o (no GPU needed):
python tools/sass_inventory.py > profiles/<round>_sass_inventory.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "VIADDMNMX", "VIMNMX", "VABSDIFF", "LEA", "IMAD", "PRMT", "LOP3", "SHF", "LDS", "STS", "LDG", "STG",
        "SHFL", "I2F", "DADD"]


def main():
    obj = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "aad_b200" / "csrc" / "build" / "aad_kernels.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("| kernel | SASS instructions | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for mangled, name in zip(counts, names):
        name = re.sub(r"\(anonymous namespace\)::|^void ", "", name).split("(")[0]
        if name.startswith("aad_"):
            c = counts[mangled]
            print(f"| `{name}` | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")


if __name__ == "__main__":
    main()
