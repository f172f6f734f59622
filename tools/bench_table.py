"""gpurun_out/bench_<config>.json -> markdown table rows + copies under profiles/r02_bench_<config>.json"""
import json, os, shutil, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
rows = []
for c in ("masking", "contextpred", "bio_supervised", "gcn", "gat", "graphsage"):
    p = os.path.join(ROOT, "gpurun_out", "bench_%s.json" % c)
    if not os.path.exists(p):
        continue
    d = json.loads(open(p).read().strip().splitlines()[-1])
    shutil.copyfile(p, os.path.join(ROOT, "profiles", "r02_bench_%s.json" % c))
    cb = d.get("cpu_baseline") or {}
    rows.append("| %s | %s | %.3f (median %.3f) | %s | %s | %d | %.3f | %.3f |" % (
        c, format(round(d["value"]), ","), d["ms_per_step"], d["detail"]["per_step_ms"]["median"], format(round(d["e2e"]["value"]), ","),
        ("%.1f (%d, %s)" % (cb["value"], cb["cores"], cb["kind"])) if cb else "—", d["gpu_launches"] // d["steps"],
        d["roofline"]["frac"], d["roofline_gather"]["frac"]))
print("| config | graphs/s (resident) | ms/step | graphs/s (e2e) | CPU reference graphs/s (cores, kind) | launches/step | GEMM family frac of bf16 peak | gather frac of HBM peak |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
print("\n".join(rows))
