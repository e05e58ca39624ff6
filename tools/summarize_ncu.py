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
        if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue
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


def layers(path, out_json=None):
    """launch list taken with T2I_NVTX=1 + `ncu --nvtx --print-nvtx-rename kernel` (tools/evidence_run.sh): the GEMM-type
    launches carry their layer shape as the kernel name.  Per layer: launches, time, share, tensor-pipe active %
    (time-weighted mean of sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active) and DRAM bytes per launch."""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
    scale_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "") or 0)
        m = row["Metric Name"]
        d = per.setdefault(row["ID"], {"name": row["Kernel Name"], "us": 0.0, "tp": 0.0, "bytes": 0.0})
        if m == "gpu__time_duration.sum":
            d["us"] = v * scale_t.get(row["Metric Unit"], 1.0)
        elif m.startswith("sm__pipe_tensor_cycles_active"):
            d["tp"] = v
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            d["bytes"] += v * scale_b.get(row["Metric Unit"], 1.0)
    agg = collections.OrderedDict()
    for d in per.values():
        name = d["name"].split("/")[0] if "/" in d["name"] else d["name"].split("(")[0].replace("void ", "").replace("t2i::", "")
        a = agg.setdefault(name, {"launches": 0, "us": 0.0, "tp_us": 0.0, "bytes": 0.0})
        a["launches"] += 1
        a["us"] += d["us"]
        a["tp_us"] += d["tp"] * d["us"]
        a["bytes"] += d["bytes"]
    tot = sum(a["us"] for a in agg.values())
    print("%d launches, %.1f us of kernel time (ncu: serialised, cold cache -- compare shares, not absolutes)" % (len(per), tot))
    print("%-54s %4s %9s %6s %12s %14s" % ("layer / kernel", "n", "us", "share", "tensor-pipe%", "DRAM MB/launch"))
    res = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        tp = a["tp_us"] / a["us"] if a["us"] else 0.0
        res[k] = {"launches": a["launches"], "us": a["us"], "tensor_pipe_active_pct": tp,
                  "dram_bytes_per_launch": a["bytes"] / a["launches"]}
        print("%-54s %4d %9.1f %5.1f%% %12.1f %14.1f" % (k[:54], a["launches"], a["us"], 100 * a["us"] / tot, tp,
                                                         a["bytes"] / a["launches"] / 1e6))
    if out_json:
        fam = collections.defaultdict(lambda: {"launches": 0, "dram_bytes_total": 0.0, "us": 0.0})
        for k, a in agg.items():
            f = "conv_gemm_kernel" if k.startswith("conv_gemm") else "wgrad_gemm_kernel" if k.startswith("wgrad_gemm") else k.split("<")[0]
            fam[f]["launches"] += a["launches"]; fam[f]["dram_bytes_total"] += a["bytes"]; fam[f]["us"] += a["us"]
        out = {f: {"launches": v["launches"], "dram_bytes_total": v["dram_bytes_total"],
                   "dram_bytes_per_launch": v["dram_bytes_total"] / v["launches"],
                   "dram_gbytes_per_s_under_ncu": v["dram_bytes_total"] / (v["us"] * 1e-6) / 1e9 if v["us"] else None}
               for f, v in sorted(fam.items(), key=lambda kv: -kv[1]["dram_bytes_total"])}
        with open(out_json, "w") as f:
            json.dump({"by_kernel": out, "by_layer": res}, f, indent=1)


def traffic(path, out_json=None):
    """dram bytes (read + write) per kernel from a launch list that also collected dram__bytes_{read,write}.sum"""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.defaultdict(lambda: {"launches": set(), "bytes": 0.0, "us": 0.0})
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("t2i::", "").split("<")[0]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg[k]
        a["launches"].add(row["ID"])
        if row["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            a["bytes"] += v * scale.get(row["Metric Unit"], 1.0)
        elif row["Metric Name"] == "gpu__time_duration.sum":
            a["us"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
    res = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["bytes"]):
        n = len(a["launches"])
        res[k] = {"launches": n, "dram_bytes_total": a["bytes"], "dram_bytes_per_launch": a["bytes"] / n,
                  "dram_gbytes_per_s_under_ncu": a["bytes"] / (a["us"] * 1e-6) / 1e9 if a["us"] else None}
        print("%-34s n=%4d  dram %9.1f MB total  %8.2f MB/launch  %7.0f GB/s (serialised)" % (
            k[:34], n, a["bytes"] / 1e6, a["bytes"] / n / 1e6, res[k]["dram_gbytes_per_s_under_ncu"] or 0))
    if out_json:
        with open(out_json, "w") as f:
            json.dump(res, f, indent=1)


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
    if sys.argv[1] == "layers":
        layers(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
