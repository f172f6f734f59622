"""Which LBO/SBO convention does the MN-major no-swizzle descriptor follow?  Small structured GEMMs."""
import importlib, os, sys, ctypes
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ops = importlib.import_module("pretrain-gnns_b200.ops")
cabi = importlib.import_module("pretrain-gnns_b200._cabi")
dll = cabi.lib.load()
dev = "cuda:0"
ops.set_precision("tf32x3")
torch.set_printoptions(linewidth=200, precision=1, sci_mode=False)
for variant in (0, 1):
    print("variant", variant, "rc", dll.pgnn_debug_set_tc_variant(variant))
    M, N, K = 8, 16, 16   # dgrad: gx[M,K] = gy[M,N] @ w[N,K]
    gy = torch.zeros(M, N, device=dev); gy[:, 0] = 1.0          # picks row n=0 of w
    w = (torch.arange(N, device=dev)[:, None] * 100 + torch.arange(K, device=dev)[None, :]).float()
    gx = ops._linear_bwd_x(gy, w)
    print("dgrad row0 (expect 0..15):", gx[0].tolist())
    gy = torch.ones(M, N, device=dev)
    gx = ops._linear_bwd_x(gy, w)
    print("dgrad ones (expect", (w.sum(0)).tolist(), "):", gx[0].tolist())
    x = torch.randn(64, 16, device=dev); g2 = torch.randn(64, 8, device=dev)
    gw, gb = ops._linear_bwd_w(g2, x)
    ref = g2.double().t() @ x.double()
    print("wgrad relerr", ((gw.double() - ref).abs().max() / ref.abs().max()).item())
    for (m, n, k) in [(5986, 600, 300), (5986, 300, 600)]:
        a = torch.randn(m, n, device=dev); ww = torch.randn(n, k, device=dev) * 0.1; xx = torch.randn(m, k, device=dev)
        r1 = a.double() @ ww.double(); r2 = a.double().t() @ xx.double()
        o1 = ops._linear_bwd_x(a, ww); o2, _ = ops._linear_bwd_w(a, xx)
        print(m, n, k, "dgrad", ((o1.double() - r1).abs().max() / r1.abs().max()).item(), "wgrad", ((o2.double() - r2).abs().max() / r2.abs().max()).item())
