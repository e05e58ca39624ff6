#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output (SASS view): per kernel, instructions executed, stall samples, and the
SASS lines with the most samples (with their dominant stall reasons).  usage: ncu_top.py file.csv [n_lines]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for si, s0 in enumerate(starts):
        end = starts[si + 1] if si + 1 < len(starts) else len(rows)
        hdr = rows[s0 + 1]
        body = [r for r in rows[s0 + 2:end] if len(r) == len(hdr)]
        col = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_")]
        inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
        samp = sum(int(r[col["# Samples"]] or 0) for r in body)
        print("== %s: %d SASS lines, %d warp-instructions, %d samples" % (rows[s0][1][:90], len(body), inst, samp))
        tot = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
        print("   stall totals:", ", ".join("%s %d" % (h[6:], v) for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        ranked = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top_n]
        for i in sorted(ranked):
            r = body[i]
            st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
            print("   %5d %-64s smp %5s exec %8s  %s" % (i, r[col["Source"]].strip()[:64], r[col["# Samples"]],
                                                        r[col["Instructions Executed"]],
                                                        " ".join("%s:%d" % (n, v) for v, n in st if v)))


if __name__ == "__main__":
    main()
