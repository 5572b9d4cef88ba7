#!/usr/bin/env python
"""Turns ncu outputs into the tracked summaries under profiles/.

  python profiles/summarize_ncu.py launches <launches.csv> <out.md> "<command>"
  python profiles/summarize_ncu.py full <report.ncu-rep> <out.md> <traffic.json> "<command>"

`launches`: per-kernel totals of a `--metrics gpu__time_duration.sum` launch list.
`full`: one row per captured launch of a `--set full` report (read with `ncu -i ... --page raw --csv`)
plus traffic.json = {kernel: dram bytes per launch} that bench.py quotes as `roofline.traffic`.
"""
import collections
import csv
import io
import json
import subprocess
import sys


def launches(src, out, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault((row["Kernel Name"].split("(")[0], row["Grid Size"], row["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"Command (under gpurun): `{cmd}`\n(cold-cache, serialised per-launch times: compare shares, not absolutes)\n\n```\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[0]:42s} {k[1]:>14s}{k[2]:>14s} n={a[0]:4d} total={a[1]:9.1f}us avg={a[1] / a[0]:8.2f}us {100 * a[1] / tot:5.1f}%\n")
        f.write(f"total us {tot:.1f}\n```\n")


COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM%"),
        ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp-insts"), ("sm__inst_executed_pipe_tensor.sum", "tensor-insts"),
        ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall-barrier")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(rep, out, traffic_json, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    traffic = collections.OrderedDict()
    with open(out, "w") as f:
        f.write(f"Command (under gpurun): `{cmd}`\n\n| kernel | grid x block | " + " | ".join(c[1] for c in COLS if c[0] in ix) + " |\n")
        f.write("|---|---|" + "---|" * sum(c[0] in ix for c in COLS) + "\n")
        for r in body:
            name = r[ix["Kernel Name"]].split("(")[0]
            cells = []
            for m, short in COLS:
                if m not in ix:
                    continue
                v, u = r[ix[m]], units[ix[m]]
                if short.startswith("dram"):
                    cells.append(f"{to_bytes(v, u) / 1e6:.3f} MB")
                elif short == "us":
                    t_us = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    cells.append(f"{t_us:.2f}")
                else:
                    cells.append(f"{float(v.replace(',', '')):.4g}" if v else "-")
            f.write(f"| {name} | {r[ix['Grid Size']]} x {r[ix['Block Size']]} | " + " | ".join(cells) + " |\n")
            t = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            traffic.setdefault(name, []).append(t)
    json.dump({k: max(v) for k, v in traffic.items()}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
