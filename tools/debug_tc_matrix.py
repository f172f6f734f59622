"""Run each (a_major, b_major, variant, shape) tcgen05 GEMM case in its own subprocess with a short timeout."""
import subprocess, sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
CASE = r'''
import importlib, os, sys, ctypes, torch
sys.path.insert(0, os.path.abspath(os.path.join(%r, "..")))
cabi = importlib.import_module("pretrain-gnns_b200._cabi")
dll = cabi.lib.load()
a_kc, b_kc, variant, M, N, K, bn = [int(v) for v in sys.argv[1:8]]
dll.pgnn_debug_set_tc_variant(variant)
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
Am = torch.randn(M, K, generator=g); Bm = torch.randn(N, K, generator=g)
ref = Am.double() @ Bm.double().t()
A = (Am if a_kc else Am.t().contiguous()).to(dev)   # a_kc: [M,K] else stored [K,M]
B = (Bm if b_kc else Bm.t().contiguous()).to(dev)
C = torch.full((M, N), float("nan"), device=dev)
f = dll.pgnn_debug_tc_gemm
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_int]*3 + [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_int]*3 + [ctypes.c_void_p]
rc = f(a_kc, b_kc, bn, A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), N, M, N, K, None)
torch.cuda.synchronize()
err = ((C.double().cpu() - ref).abs().max() / ref.abs().max()).item()
# which rows / cols are right?
good = ((C.double().cpu() - ref).abs() < 1e-3 * ref.abs().max())
print("rc", rc, "relerr %%.2e" %% err, "good rows", int(good.all(1).sum()), "/", M, "good cols", int(good.all(0).sum()), "/", N)
''' % HERE
cases = []
for (a, b) in [(0, 1), (1, 0), (0, 0)]:
    for variant in (0, 1):
        for (M, N, K, bn) in [(8, 16, 8, 64), (128, 64, 64, 64), (128, 224, 96, 224), (300, 600, 5986, 224)]:
            cases.append((a, b, variant, M, N, K, bn))
for c in cases:
    try:
        r = subprocess.run([sys.executable, "-c", CASE] + [str(v) for v in c], capture_output=True, text=True, timeout=25)
        out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
        if r.returncode != 0:
            out += " | rc=%d %s" % (r.returncode, (r.stderr.strip().splitlines() or [""])[-1][:150])
    except subprocess.TimeoutExpired:
        out = "TIMEOUT"
    print("a_kc=%d b_kc=%d variant=%d M=%d N=%d K=%d bn=%d ->" % c, out, flush=True)
