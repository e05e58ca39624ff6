#!/usr/bin/env python
"""Summarise ncu outputs for profiles/:  launch list CSV (gpu__time_duration) -> per-kernel table;
full-capture .ncu-rep -> key metrics per captured launch (needs `ncu` on PATH, no GPU)."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum.per_second"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("%d launches, %.1f us of kernel time (ncu: serialised, cold cache -- compare shares, not absolutes)" % (
        sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s n=%4d  us=%9.1f  %5.1f%%" % (k[:58], v[0], v[1], 100 * v[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d.get("Kernel Name", "?"), " grid", d.get("launch__grid_size"))
        for k in KEYS:
            if k in d:
                print("   %-78s %s %s" % (k, d[k], units[hdr.index(k)]))


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
