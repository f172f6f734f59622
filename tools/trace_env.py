"""Phase envelope over every CTA of the TMA GEMM (PGNN_TRACE_ALL build; PGNN_LIB points at it): for each stamp the earliest
and latest CTA, relative to the earliest kernel entry.  Shapes: the six GEMMs of one chem GIN layer at B = 256."""
import ctypes, importlib, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ops = importlib.import_module("pretrain-gnns_b200.ops")
cabi = importlib.import_module("pretrain-gnns_b200._cabi")
dll = cabi.lib.load()
ops.set_precision("tf32x3")
dev = "cuda:0"
names = ["start", "alloc+init done", "first block converted", "last block converted", "mma first stage", "mma last stage",
         "acc complete", "epilogue done", "all done (post sync)", "tmem->smem done"]
fn = dll.pgnn_debug_tma_trace_env
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def report(tag, launch):
    for rep in range(2):
        launch()
    torch.cuda.synchronize()
    flush.zero_(); torch.cuda.synchronize()
    fn(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); launch(); b.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    fn(buf, 0)
    t = list(buf)
    t0 = t[0]
    print("%s: event time %.2f us" % (tag, a.elapsed_time(b) * 1e3))
    for i, n in enumerate(names):
        if t[i] != 2 ** 64 - 1 and t[16 + i]:
            print("    %-24s earliest +%6.2f us   latest +%6.2f us" % (n, (t[i] - t0) / 1e3, (t[16 + i] - t0) / 1e3))


M = 5986
x3, x6 = torch.randn(M, 300, device=dev), torch.randn(M, 600, device=dev)
w1, w2 = torch.randn(600, 300, device=dev) * 0.05, torch.randn(300, 600, device=dev) * 0.05
b6, b3 = torch.zeros(600, device=dev), torch.zeros(300, device=dev)
report("GEMM1 fwd [M,300]x[300,600]", lambda: ops._linear_fwd(x3, w1, b6, True))
report("GEMM2 fwd [M,600]x[600,300]", lambda: ops._linear_fwd(x6, w2, b3, False))
report("dgrad2 [M,300]x[300->600] (MN-major weight)", lambda: ops._linear_bwd_x(x3, w2, mask=x6))
report("wgrad2 gw[300,600]", lambda: ops._linear_bwd_w(x3, x6))
report("wgrad1 gw[600,300]", lambda: ops._linear_bwd_w(x6, x3))
