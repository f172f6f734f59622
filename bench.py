#!/usr/bin/env python
"""bench.py — graphs/sec through the 5-layer, 300-dim chem GIN forward+backward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE configs[1]): the train() body of chem/pretrain_masking.py:46-70 on synthetic
ZINC-shaped batches of 256 graphs — GNN(5, 300, JK="last", drop_ratio=0, "gin") in train mode,
`linear_pred_atoms(node_rep[masked_atom_indices])`, cross-entropy on fp64 logits, `loss.backward()`.
Optimizer steps are excluded; with N > 1 every rank draws its own batch (weak scaling,
graphs sharded by rank) and the step includes ONE all-reduce of the flat fp32 gradient buffer.

One JSON line is printed by rank 0.  `value` times K steps with the batch already resident in HBM
(graph bucketing included: every step sees a different batch); `e2e` times the same K steps from pinned
host memory (x / edge_index / edge_attr / mask indices / labels packed into one pinned buffer per batch, one
asynchronous H2D copy per step inside the timed region with the next batch prefetched under the running step, loss
read back synchronously every step).  L2 is flushed between timed steps (256 MiB memset outside the per-step event pairs).
`--impl reference` times the CPU oracle port of the reference's model.py on the host cores.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 256
NUM_LAYER, EMB = 5, 300
NUM_DISTINCT_BATCHES = 8
METRIC = "graphs/sec 5-layer GIN-300 fwd+bwd on ZINC-shaped batches"
GATHER_DRAM_BYTES_NCU = 7281664  # k_aggregate_bwd, one `ncu --set full` capture (profiles/r01_mid_kernels_ncu.md)
GEMM1_DRAM_BYTES_NCU = 7956992  # dram__bytes_read.sum + dram__bytes_write.sum of the B=256 GEMM1 launch (profiles/r01_gemm_tma_ncu.md)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled every 50 ms while the timed region runs.

    In-process NVML (nvidia_ml_py) in a daemon thread: spawning `nvidia-smi -lms` next to the timed region made the
    first end-to-end measurement on a fresh box ~2x slower (its start-up contends for the driver while the step is
    CPU-launch-bound); nvidia-smi is only the fallback when NVML cannot be imported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.stop = index, [], None, False
        self.sm, self.mx, self.reasons, self.t = [], [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass

    def __exit__(self, *a):
        self.stop = True
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
        if self.t is not None:
            self.t.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(self.mx), "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def make_batches(syn, rank, count):
    out = []
    for i in range(count):
        seed = 2000 + 1000 * rank + i  # config-id*1000 + rank + batch index
        b = syn.mask_atoms(syn.zinc_batch(BATCH, seed), seed)
        out.append({k: b[k] for k in ("x", "edge_index", "edge_attr", "masked_atom_indices")} |
                   {"labels": b["mask_node_label"][:, 0].contiguous()})
    return out


# ---------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle port of chem/model.py + the masking head, all host threads
# ---------------------------------------------------------------------------------------------------
def cpu_oracle_run(steps, warmup, threads=None, batches=None, one_thread=False):
    from oracle import gnn_oracle as O
    syn = importlib.import_module("pretrain-gnns_b200.synthetic")
    cores = os.cpu_count() or 1
    P = O.leaf_params(O.make_params("chem", "gin", NUM_LAYER, EMB, seed=1, randomize_bn=False))
    g = torch.Generator().manual_seed(5)
    W = (torch.randn(119, EMB, generator=g) * 0.05).requires_grad_(True)
    bvec = torch.zeros(119, requires_grad=True)
    batches = batches or make_batches(syn, 0, 2)

    def step(b):
        for v in list(P.values()) + [W, bvec]:
            if v.requires_grad:
                v.grad = None
        rep = O.chem_gnn(P, b["x"], b["edge_index"], b["edge_attr"], NUM_LAYER, "gin", True)
        loss, _ = O.masking_loss(rep, b["masked_atom_indices"], b["labels"], W, bvec)
        loss.backward()
        return float(loss.detach())

    if threads is None:
        # "all the host threads it can use": these ops are small, so past a point more threads only add
        # synchronisation cost; probe powers of two up to the core count and keep the fastest
        best = (float("inf"), 1)
        cand = sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= min(cores, 64)} | {min(cores, 8)})  # 128 threads: 25 graphs/s (measured), not probed
        for c in cand:
            torch.set_num_threads(c)
            step(batches[0])
            t0 = time.perf_counter()
            step(batches[0])
            best = min(best, (time.perf_counter() - t0, c))
        threads = best[1]
    torch.set_num_threads(threads)
    for i in range(warmup):
        step(batches[i % len(batches)])
    ts = []
    for i in range(steps):
        t0 = time.perf_counter()
        step(batches[i % len(batches)])
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    one = None
    if one_thread:  # "also report a 1-thread figure"
        torch.set_num_threads(1)
        step(batches[0])
        t0 = time.perf_counter()
        for i in range(2):
            step(batches[i % len(batches)])
        one = BATCH * 2 / (time.perf_counter() - t0)
        torch.set_num_threads(threads)
    return dict(value=BATCH * steps / total, ms_per_step=1e3 * total / steps, cores=threads, value_1thread=one,
                sample="%d fwd+bwd steps (after %d warm-up) of the B=%d masking batch (oracle port of chem/model.py, torch CPU, best of the "
                       "probed thread counts = %d of %d host cores)" % (steps, warmup, BATCH, threads, cores))


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_oracle_run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "chem pretrain_masking 5-layer GIN emb_dim=300 batch_size=256 (BASELINE configs[1])",
                       "global_batch": BATCH, "note": "CPU oracle port; real torch_geometric 1.0.3 is not installable"},
            "cpu_baseline": {"value": r["value"], "unit": "graphs/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    syn = importlib.import_module("pretrain-gnns_b200.synthetic")
    chem = importlib.import_module("pretrain-gnns_b200.chem.model")
    ops = importlib.import_module("pretrain-gnns_b200.ops")
    cabi = importlib.import_module("pretrain-gnns_b200._cabi")
    pdist = importlib.import_module("pretrain-gnns_b200.dist")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback (use --impl reference)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if args.precision:
        ops.set_precision(args.precision)
    torch.manual_seed(0)
    model = chem.GNN(NUM_LAYER, EMB, JK="last", drop_ratio=0, gnn_type="gin").to(dev).train()
    head = torch.nn.Linear(EMB, 119).to(dev)
    params = list(model.parameters()) + list(head.parameters())
    reducer = pdist.GradAllReducer(params, flat_sources=[pdist.encoder_flat_source(model)]) if world > 1 else None

    host = make_batches(syn, rank, NUM_DISTINCT_BATCHES)
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(b):
        for p in params:
            p.grad = None
        rep = model(b["x"], b["edge_index"], b["edge_attr"])
        # linear_pred_atoms(node_rep[masked_atom_indices]) + CrossEntropyLoss on .double() logits (reference :51-52)
        loss, _ = ops.masked_atom_loss(rep, b["masked_atom_indices"], b["labels"], head.weight, head.bias)
        loss.backward()
        if reducer is not None:
            reducer.all_reduce_mean()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # end to end: every batch is ONE pinned buffer and ONE asynchronous copy on a side stream (data.BatchStager); the copy of
    # batch i+1 is issued right after step i has been enqueued, so it runs under step i's kernels.  Each of the K timed steps
    # still contains exactly one host->device copy (the first step's own, then each step's prefetch of the next) and the
    # synchronous read of its loss, as chem/pretrain_masking.py:76 does.
    pdata = importlib.import_module("pretrain-gnns_b200.data")
    stager = pdata.BatchStager(dev)
    packed = [stager.pack(b) for b in host]
    h2d_bytes = packed[0].nbytes  # what one step copies (the five tensors, each padded to 16 bytes)

    pinned = None

    def simple_fetch(i):  # fallback transport: one .to() per tensor from pinned memory, in front of the step
        nonlocal pinned
        if pinned is None:
            pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        return {k: v.to(dev, non_blocking=True) for k, v in pinned[i % len(pinned)].items()}

    def timed(e2e, staged=True):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()  # ranks finish their (CPU-side) setup seconds apart; start the first exchange together
        for i in range(args.warmup):
            if not e2e:
                b = resident[i % len(resident)]
            else:
                b = stager.take(stager.submit(packed[i % len(packed)])) if staged else simple_fetch(i)
            step(b).item()
        barrier()
        n0 = cabi.lib.pgnn_kernel_launch_count()
        t0 = time.perf_counter()
        ticket = None
        for i in range(args.steps):
            flush.zero_()  # L2 flush, outside the per-step event pair
            ev[i][0].record()
            if e2e and staged:
                if ticket is None:
                    ticket = stager.submit(packed[i % len(packed)])
                loss = step(stager.take(ticket))
                ticket = stager.submit(packed[(i + 1) % len(packed)]) if i + 1 < args.steps else None
                loss.item()   # D2H read of the loss every step
            elif e2e:
                step(simple_fetch(i)).item()
            else:
                step(resident[i % len(resident)])
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - t0
        launches = cabi.lib.pgnn_kernel_launch_count() - n0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), launches, wall

    with ClockSampler(local_rank) as clocks:
        ms_dev, launches, wall_dev = timed(False)
        try:
            ms_e2e, _, wall_e2e = timed(True)
            e2e_transport = "one pinned buffer + one async copy per batch on a side stream, next batch prefetched under the step (data.BatchStager)"
        except Exception as e:  # (after a sticky CUDA error the fallback fails too and the run ends with that error)
            print("[bench] staged end-to-end path failed (%s: %s); per-tensor copies instead" % (type(e).__name__, e), file=sys.stderr, flush=True)
            ms_e2e, _, wall_e2e = timed(True, staged=False)
            e2e_transport = "one .to(device) per tensor from pinned memory in front of each step"
    graphs = BATCH * world * args.steps

    roof, roof_gather = kernel_rooflines(ops, cabi, resident[0], dev) if rank == 0 else (None, None)
    if rank == 0 and world == 1:
        try:  # per-kernel durations INSIDE the step (library timing mode: CUDA event pairs around every launch, warm L2)
            in_step_rooflines(cabi, step, resident, flush, roof, roof_gather)
        except Exception as e:  # diagnostic only: the isolated timings above stand
            print("[bench] in-step kernel timing skipped: %s: %s" % (type(e).__name__, e), file=sys.stderr, flush=True)
    cpu = cpu_oracle_run(10, 3, batches=host[:2], one_thread=True) if rank == 0 and world == 1 and not args.no_cpu_baseline else None  # N=1 only
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": graphs / (ms_dev * 1e-3), "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if ops.get_precision() == "fp32" else "tf32x3", "data": "synthetic",
        "config": {"workload": "chem pretrain_masking 5-layer GIN emb_dim=300 batch_size=256 (BASELINE configs[1])",
                   "global_batch": BATCH * world, "per_gpu_batch": BATCH, "parallelism": "dp%d" % world,
                   "nodes_per_batch": int(host[0]["x"].shape[0]), "edges_per_batch": int(host[0]["edge_index"].shape[1]),
                   "distinct_batches": NUM_DISTINCT_BATCHES, "l2": "flushed between timed steps (256 MiB memset)",
                   "optimizer_step": "excluded (SURVEY 8(d))", "gemm_precision": ops.get_precision(),
                   "grad_allreduce": (reducer.backend if reducer is not None else "none (1 GPU)"),
                   "wall_ms_per_step_incl_flush": 1e3 * wall_dev / args.steps},
        "e2e": {"value": graphs / (ms_e2e * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                "ms_per_step": ms_e2e / args.steps, "transport": e2e_transport},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roof, "roofline_gather": roof_gather,
        "cpu_baseline": None if cpu is None else {"value": cpu["value"], "unit": "graphs/s", "cores": cpu["cores"], "kind": "port",
                                                  "sample": cpu["sample"], "value_1thread": cpu["value_1thread"]},
    }
    print(json.dumps(line), flush=True)


def in_step_rooflines(cabi, step, resident, flush, roof, roof_gather, steps=6):
    """Re-run a few steps with the library's timing mode on and restate the two rooflines with each kernel's average
    duration inside the real step (its operands where the previous kernel left them) instead of the isolated launch."""
    import ctypes
    lib = cabi.lib
    torch.cuda.synchronize()
    lib.pgnn_profile_read(None, 0)  # drop anything recorded earlier
    lib.pgnn_profile_enable(1)
    try:
        for i in range(steps):
            flush.zero_()
            step(resident[i % len(resident)])
        torch.cuda.synchronize()
    finally:
        lib.pgnn_profile_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib.pgnn_profile_read(buf, len(buf))
    rows = {}
    for line in buf.value.decode(errors="replace").splitlines():
        name, cnt, us = line.rsplit("\t", 2)
        rows[name] = (int(cnt), float(us))
    if n <= 0 or not rows:
        raise RuntimeError("no launches recorded")
    total = sum(us for _, us in rows.values())

    def pick(base, *variants):  # kernel names come back mangled; accept the demangled spelling as well
        c = [(cnt, us) for name, (cnt, us) in rows.items() if base in name and (not variants or any(v in name for v in variants))]
        return (sum(a for a, _ in c), sum(b for _, b in c)) if c else (0, 0.0)

    # GEMM1 forward and dgrad2 share the instantiation <K-major, K-major, 224> and the shape [N,300]x[300,600]
    cnt, us = pick("k_gemm_3xtf32_tma", "Lb0ELb0ELi224E", "<false, false, 224", "<0, 0, 224", "<(bool)0, (bool)0, 224")
    if cnt:
        avg = us / cnt
        roof["isolated_us_per_launch"] = roof["us_per_launch"]
        roof["us_per_launch"] = avg
        roof["achieved"] = roof["achieved"] * roof["isolated_us_per_launch"] / avg
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["timing"] = ("average of %d launches inside %d training steps (CUDA event pairs around each launch, library timing "
                          "mode); GEMM1 forward and the dgrad of GEMM2 share this instantiation and shape" % (cnt, steps))
        roof["share_of_step"] = sum(b for nm, (a, b) in rows.items() if "k_gemm_3xtf32" in nm) / total
    cnt, us = pick("k_aggregate_fwd")
    if cnt:
        avg = us / cnt
        roof_gather["isolated_us_per_launch"] = roof_gather["us_per_launch"]
        roof_gather["us_per_launch"] = avg
        roof_gather["achieved"] = roof_gather["algorithmic_bytes"] / (avg * 1e-6) / 1e9
        roof_gather["frac"] = roof_gather["achieved"] / roof_gather["peak"]
        roof_gather["timing"] = "average of %d launches inside %d training steps; the activations are L2-resident there" % (cnt, steps)
        roof_gather["share_of_step"] = sum(b for nm, (a, b) in rows.items() if "k_aggregate" in nm) / total
    import re

    def short(nm):  # "_ZN..17k_gemm_3xtf32_tmaILb0ELb0ELi224ELb0EEEv..." -> "k_gemm_3xtf32_tma<0,0,224,0>"
        m = re.search(r"(k_[A-Za-z0-9_]+?)(I(?:L[bi]\d+E)+E)?(?:Ev|E?P|$)", nm)
        if not m:
            return nm[:60]
        targs = re.findall(r"L[bi](\d+)E", m.group(2) or "")
        return m.group(1) + ("<" + ",".join(targs) + ">" if targs else "")

    agg = {}
    for nm, (_, us) in rows.items():
        agg[short(nm)] = agg.get(short(nm), 0.0) + us
    roof["step_kernels_us"] = {k: round(v / steps, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]}
    roof["step_sum_us_serialised"] = round(total / steps, 1)  # library kernels only, each bracketed by events (no PDL overlap)


def kernel_rooflines(ops, cabi, b, dev):
    """Isolated timings (CUDA events on the launch stream, L2 flushed before each launch) of the two kernels
    the step is made of: the MLP GEMM (dominant, tensor/FMA-bound) and the neighbour gather (HBM/L2-bound)."""
    pk = peaks()
    n, e = int(b["x"].shape[0]), int(b["edge_index"].shape[1])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = ops.Graph(b["edge_index"], n)
    S = g.summary("chem", ops.AGG_SUM, b["edge_attr"])
    x = torch.randn(n, EMB, device=dev)
    T = torch.randn(9, EMB, device=dev)
    w1, b1 = torch.randn(2 * EMB, EMB, device=dev) * 0.05, torch.zeros(2 * EMB, device=dev)

    def avg_ms(fn, reps=20):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); c.record(); c.synchronize()
            tot += a.elapsed_time(c)
        return tot / reps

    with torch.no_grad():
        t_gather = avg_ms(lambda: ops.aggregate(x, T, g, S, ops.AGG_SUM))
        a = ops.aggregate(x, T, g, S, ops.AGG_SUM)
        t_gemm = avg_ms(lambda: ops._linear_fwd(a, w1, b1, True))
    gbytes = 4 * EMB * (e + 2 * n)              # rows read (E+N) + rows written N
    gflop = 2.0 * n * EMB * 2 * EMB             # GEMM1 of the MLP: [N,300] x [300,600]
    mode = ops.get_precision()
    ach = gflop / (t_gemm * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "MLP GEMM1 [N,300]x[300,600] + bias + ReLU (%s; k_gemm_3xtf32_tma<0,0,224>)" % mode,
            "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": ach / pk["tensor"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at these shapes, one `ncu --set full` capture
            # (profiles/r01_gemm_tma_ncu.md): 7.96 MB read + 0 B written inside the kernel window; algorithmic 7.18 + 0.72 MB
            "traffic": GEMM1_DRAM_BYTES_NCU,
            "peak_source": pk["src"] + " cuBLAS bf16 dense burst (MEASURED_PEAKS.json)",
            "note": "fp32-equivalent flops; 3xTF32 spends 3 tf32 MACs per fp32 MAC and dense tf32 is half of bf16, so the "
                    "ceiling of this scheme is peak/6 = %.0f TFLOP/s (achieved/ceiling = %.3f)" % (pk["tensor"] / 6, ach / (pk["tensor"] / 6)),
            "us_per_launch": t_gemm * 1e3}
    roof_g = {"bound": "hbm", "kernel": "k_aggregate_fwd (gather + segment sum, one layer pass)", "achieved": gbytes / (t_gather * 1e-3) / 1e9,
              "peak": pk["hbm"], "unit": "GB/s", "frac": gbytes / (t_gather * 1e-3) / 1e9 / pk["hbm"],
              # dram__bytes_read + write of the transpose-graph twin k_aggregate_bwd at these shapes (profiles/r01_mid_kernels_ncu.md):
              # the matrix crosses DRAM once (7.28 MB) and the (E+N) row reads are served by L2 (39.4 MB of lts sectors)
              "traffic": GATHER_DRAM_BYTES_NCU,
              "peak_source": pk["src"], "us_per_launch": t_gather * 1e3, "algorithmic_bytes": gbytes}
    return roof, roof_g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "tf32x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.steps > 500:
            args.steps = 500  # bounded sample: ~0.1 s per CPU step at the best thread count
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
