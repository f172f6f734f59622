"""Decompose the TMA GEMM's time by role, on the GPU box:  python tools/ubench_gemm.py [build|run]

`build` (no GPU needed) compiles dense_tma.cu five times with the PGNN_UB_* switches and links each against the product
objects into tools/_ub/libpgnn_<variant>.so.  `run` executes tools/check_tc.py quick + tools/trace_tc.py once per variant
(PGNN_LIB selects the library) and prints the per-block main-loop time of each: full, no TMA (convert + MMA), no convert
(TMA + MMA), no MMA (TMA + convert), no epilogue.  Results of the cut-down variants are garbage by construction."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "pretrain-gnns_b200")
OUT = os.path.join(ROOT, "tools", "_ub")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
VARIANTS = {"full": [], "no_tma": ["-DPGNN_UB_NO_TMA"], "no_convert": ["-DPGNN_UB_NO_CONVERT"], "no_mma": ["-DPGNN_UB_NO_MMA"],
            "no_epilogue": ["-DPGNN_UB_NO_EPILOGUE"], "mma_only": ["-DPGNN_UB_NO_TMA", "-DPGNN_UB_NO_CONVERT"],
            "trace_all": ["-DPGNN_TRACE_ALL"]}
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "--expt-relaxed-constexpr"]


def build():
    subprocess.check_call([sys.executable, os.path.join(PKG, "build.py")])
    os.makedirs(OUT, exist_ok=True)
    objs = [o for o in glob.glob(os.path.join(PKG, "csrc", "_obj", "*.o")) if not o.endswith("dense_tma.o")]
    for name, defs in VARIANTS.items():
        obj = os.path.join(OUT, "dense_tma_%s.o" % name)
        subprocess.check_call([NVCC] + FLAGS + defs + ["-c", os.path.join(PKG, "csrc", "dense_tma.cu"), "-o", obj])
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(OUT, "libpgnn_%s.so" % name)]
                              + objs + [obj, "-cudart", "static"])
        print("built", name)


def run_env():
    """Envelope over ALL CTAs of the full kernel (trace_all build): earliest / latest stamp of every phase."""
    env = dict(os.environ, PGNN_LIB=os.path.join(OUT, "libpgnn_trace_all.so"))
    tr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "trace_env.py")], env=env, capture_output=True, text=True, timeout=180)
    print(tr.stdout)
    if tr.returncode:
        print("trace_env failed:", tr.stderr[-400:])


def run():
    run_env()
    for name in VARIANTS:
        if name == "trace_all":
            continue
        env = dict(os.environ, PGNN_LIB=os.path.join(OUT, "libpgnn_%s.so" % name))
        tr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "trace_tc.py")], env=env, capture_output=True, text=True, timeout=120)
        for line in tr.stdout.splitlines():
            m = re.match(r"fwd M=(\d+) N=(\d+) K=(\d+):", line)
            if not m:
                continue
            t = dict((k.strip(), float(v)) for k, v in re.findall(r"([a-z0-9:>\- +]+?) \+([0-9.]+)us", line.split(":", 1)[1]))
            nkb = (int(m.group(3)) + 31) // 32
            a, b = t.get("mma: first stage ready"), t.get("mma: last stage ready")
            per = (b - a) / max(nkb - 1, 1) if a is not None and b is not None else float("nan")
            print("%-12s M=%5s N=%4s K=%4s  per 32-deep block %.3f us | acc complete +%.2f | all done +%.2f"
                  % (name, m.group(1), m.group(2), m.group(3), per, t.get("acc complete", float("nan")), t.get("all epilogue done", float("nan"))),
                  flush=True)
        if tr.returncode:
            print(name, "trace failed:", tr.stderr[-300:])


if __name__ == "__main__":
    (build if (len(sys.argv) < 2 or sys.argv[1] == "build") else run)()
