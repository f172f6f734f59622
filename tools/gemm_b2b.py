"""Back-to-back launch throughput of the GEMM shapes of one chem GIN layer (warm L2, the way the step runs them): average device
time per launch over a train of launches between one CUDA-event pair.  PGNN_LIB / PGNN_TMA_STORE select the build / epilogue."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ops = importlib.import_module("pretrain-gnns_b200.ops")
ops.set_precision("tf32x3")
dev = "cuda:0"
M = int(os.environ.get("M", 5986))
x3, x6 = torch.randn(M, 300, device=dev), torch.randn(M, 600, device=dev).relu()
w1, w2 = torch.randn(600, 300, device=dev) * 0.05, torch.randn(300, 600, device=dev) * 0.05
b6, b3 = torch.zeros(600, device=dev), torch.zeros(300, device=dev)


def train(tag, fn, reps=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("%-44s %7.2f us per launch" % (tag, a.elapsed_time(b) * 1e3 / reps), flush=True)


train("GEMM1 fwd [M,300]x[300,600] +bias+relu", lambda: ops._linear_fwd(x3, w1, b6, True))
train("GEMM2 fwd [M,600]x[600,300] +bias", lambda: ops._linear_fwd(x6, w2, b3, False))
train("dgrad2 [M,300]x[300->600] masked", lambda: ops._linear_bwd_x(x3, w2, mask=x6))
train("dgrad1 [M,600]x[600->300]", lambda: ops._linear_bwd_x(x6, w1))
train("wgrad2 gw[300,600]", lambda: ops._linear_bwd_w(x3, x6))
train("wgrad1 gw[600,300]", lambda: ops._linear_bwd_w(x6, x3))
