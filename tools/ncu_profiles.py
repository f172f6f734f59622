"""Turn the ncu captures a gpurun call brought back into the committed summaries under profiles/.

    python tools/ncu_profiles.py launches gpurun_out/launches_masking.csv profiles/r02_launches_masking.md --steps 2
    python tools/ncu_profiles.py full gpurun_out/prof_masking.ncu-rep profiles/r02_kernels_masking.md --config masking

`launches`: per-kernel totals / shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
`full`    : per-kernel table of an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`): duration, DRAM
            bytes read + written, L2 (lts) bytes, tensor-pipe and warp activity; the GEMM and gather rows also go to
            profiles/traffic.json, which bench.py reads for `roofline.traffic` (no constants typed into bench.py).
"""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:70]


def read_csv_rows(text):
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def launches(args):
    rows = read_csv_rows(open(args.src).read())
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    out = ["# %s" % args.title, "", args.note, "",
           "| kernel | launches | avg µs | total µs | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.2f | %.1f | %.1f%% |" % (k, n, us / n, us, 100 * us / total))
    out.append("")
    out.append("Total %.1f µs over %d launches (%d step(s): %.1f µs per step, cold-cache and serialised by ncu)."
               % (total, sum(n for n, _ in agg.values()), args.steps, total / max(args.steps, 1)))
    open(args.dst, "w").write("\n".join(out) + "\n")
    print("wrote", args.dst)


METRICS = {"gpu__time_duration.sum": "dur", "dram__bytes_read.sum": "dr", "dram__bytes_write.sum": "dw", "lts__t_sectors.sum": "l2",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor",
           "sm__warps_active.avg.pct_of_peak_sustained_active": "warps", "launch__registers_per_thread": "regs",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
           "lts__t_sectors_srcunit_tex_op_read.sum": "l2_rd_sectors", "smsp__inst_executed.sum": "inst"}
SCALE = {"sector": 32, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}


def full(args):
    text = subprocess.run(["ncu", "-i", args.src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    per = collections.OrderedDict()
    for r in rows[2:]:
        k = short(r[col["Kernel Name"]])
        d = per.setdefault(k, collections.defaultdict(list))
        for m, key in METRICS.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                d[key].append(float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1))
    out = ["# %s" % args.title, "", args.note, "",
           "| kernel | launches | µs | DRAM read MB | DRAM write MB | L2 sectors MB | L2 throughput % | tensor pipe % | issue active % | warps active % | regs |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    avg = lambda xs: sum(xs) / len(xs) if xs else float("nan")
    traffic = {}
    for k, d in per.items():
        out.append("| `%s` | %d | %.2f | %.2f | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %d |"
                   % (k, len(d["dur"]), avg(d["dur"]), avg(d["dr"]) / 1e6, avg(d["dw"]) / 1e6, avg(d["l2"]) / 1e6, avg(d["l2_pct"]), avg(d["tensor"]),
                      avg(d["issue"]), avg(d["warps"]),
                      int(avg(d["regs"])) if d["regs"] else -1))
        role = "gemm" if "k_gemm_3xtf32_tma" in k and args.gemm_pattern in k else ("gather" if k.startswith(args.gather_kernel) else None)
        if role and role not in traffic:
            traffic[role] = {"dram_bytes_per_launch": int(avg(d["dr"]) + avg(d["dw"])), "kernel": k, "launches_captured": len(d["dur"]),
                             "source": os.path.relpath(args.dst, ROOT) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    open(args.dst, "w").write("\n".join(out) + "\n")
    print("wrote", args.dst)
    if args.config and traffic:
        p = os.path.join(ROOT, "profiles", "traffic.json")
        allc = json.load(open(p)) if os.path.exists(p) else {}
        allc.setdefault(args.config, {}).update(traffic)
        json.dump(allc, open(p, "w"), indent=1, sort_keys=True)
        print("updated", p, traffic)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["launches", "full"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--note", default="")
    ap.add_argument("--config", default=None)
    ap.add_argument("--gemm-pattern", default="<0, 0, 208", help="substring selecting the roofline GEMM instantiation")
    ap.add_argument("--gather-kernel", default="k_aggregate_fwd")
    a = ap.parse_args()
    (launches if a.mode == "launches" else full)(a)
