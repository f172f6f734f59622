"""Phase timeline of CTA (0,0,0) of the tcgen05 GEMM (globaltimer stamps written by the kernel)."""
import importlib, os, sys, ctypes, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ops = importlib.import_module("pretrain-gnns_b200.ops")
cabi = importlib.import_module("pretrain-gnns_b200._cabi")
dll = cabi.lib.load()
ops.set_precision("tf32x3")
dev = "cuda:0"
names = ["start", "alloc+init done", "first block published", "last block published", "mma: first stage ready", "mma: last stage ready",
         "acc complete", "warp0 epilogue done", "all epilogue done", "warp0 tmem->smem done"]
for (M, N, K) in [(5986, 600, 300), (130, 600, 300), (5986, 300, 600)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1; b = torch.zeros(N, device=dev)
    for rep in range(3):
        y = ops._linear_fwd(x, w, b, True)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    (dll.pgnn_debug_tc_trace if os.environ.get('PGNN_NO_TMA') == '1' else dll.pgnn_debug_tma_trace)(buf)
    t = list(buf)
    print("fwd M=%d N=%d K=%d:" % (M, N, K), " | ".join("%s +%.2fus" % (n, (t[i] - t[0]) / 1e3) for i, n in enumerate(names)))
