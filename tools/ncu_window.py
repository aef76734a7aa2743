#!/usr/bin/env python3
"""Print a window of SASS instructions with sample counts and stall reasons from
`ncu -i X.ncu-rep --page source --csv`.  usage: ncu_window.py source.csv <start_index> <count> [kernel_index]"""
import csv
import sys

KEYS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_no_inst",
        "stall_dispatch", "stall_branch_resolving", "stall_selected", "stall_not_selected", "stall_barrier"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start, count = int(sys.argv[2]), int(sys.argv[3])
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    sec = secs[which]
    ix = {h: i for i, h in enumerate(sec["hdr"])}
    tot = 0
    for k, r in enumerate(sec["rows"][start:start + count]):
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        bd = " ".join(f"{key[6:]}={r[ix[key]]}" for key in KEYS if key in ix and int(r[ix[key]] or 0))
        print(f"{start + k:5d} {s:7d}  {r[ix['Source']][:60]:60s} {bd}")
    print("window samples:", tot)


if __name__ == "__main__":
    main()
