#!/usr/bin/env python
"""SASS digest of the built library (no GPU needed): per kernel, the counts of the opcodes that prove the Blackwell
paths -- UTCHMMA (tcgen05.mma, .2CTA = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tensor loads / stores),
UBLKCP (bulk copies), UTCBAR (tcgen05.commit), SYNCS (mbarrier), LDGSTS (cp.async), REDG/RED (global reductions).
    python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "text-to-image_b200", "libt2i_b200.so")
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "REDG", "ATOMG", "HMMA", "FFMA",
        "BAR.SYNC", "STS", "LDS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(t2i::\w+\)$", "", name).replace("void ", "").replace("t2i::", "").replace("(bool)", "").replace("(int)", "")
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    counts[name][k] += 1
                    break
    print("SASS digest of %s (cuobjdump -sass; opcode counts per kernel, instructions in total)" % os.path.relpath(LIB, ROOT))
    print("%-62s %6s  %s" % ("kernel", "instr", "  ".join(KEYS[:11])))
    tot = collections.Counter()
    for n, c in counts.items():
        tot.update(c)
        if not any(c[k] for k in KEYS[:10]):
            continue
        print("%-62s %6d  %s" % (n[:62], total[n], "  ".join("%*d" % (len(k), c[k]) for k in KEYS[:11])))
    print("%-62s %6d  %s" % ("ALL KERNELS (%d)" % len(counts), sum(total.values()), "  ".join("%*d" % (len(k), tot[k]) for k in KEYS[:11])))


if __name__ == "__main__":
    main()
