#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per kernel, stall-reason totals and the
hottest SASS instructions.  usage: ncu_stalls.py source.csv [top_n]"""
import csv
import sys

KEYS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_no_inst",
        "stall_dispatch", "stall_branch_resolving", "stall_selected", "stall_not_selected", "stall_barrier",
        "stall_sleep", "stall_tex", "stall_membar", "stall_drain", "stall_misc"]


def sections(path):
    rows = list(csv.reader(open(path)))
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            yield cur
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)


def main():
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    secs = list(sections(sys.argv[1]))
    for sec in secs:
        ix = {h: i for i, h in enumerate(sec["hdr"])}
        if "# Samples" not in ix:
            continue
        data = sec["rows"]
        num = lambda r, k: int(r[ix[k]] or 0) if k in ix else 0
        total = sum(num(r, "# Samples") for r in data)
        if total == 0:
            continue
        print(f"== {sec['name']}  ({len(data)} SASS instructions, {total} samples)")
        agg = {k: sum(num(r, k) for r in data) for k in KEYS}
        print("   " + "  ".join(f"{k[6:]}={100.0 * v / total:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
        for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
            s = num(r, "# Samples")
            bd = " ".join(f"{k[6:]}={num(r, k)}" for k in KEYS if num(r, k))
            print(f"   {r[ix['Address']][-5:]} {100.0 * s / total:5.2f}%  {r[ix['Source']][:58]:58s} {bd}")


if __name__ == "__main__":
    main()
