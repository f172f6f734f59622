"""Launch the tensor-core GEMMs of one GIN layer (B=256 shapes) a few times — target for ncu."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
ops = importlib.import_module("pretrain-gnns_b200.ops")
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
dev = "cuda:0"
M, D = 5986, 300
x = torch.randn(M, D, device=dev); w1 = torch.randn(2 * D, D, device=dev) * 0.05; b1 = torch.zeros(2 * D, device=dev)
w2 = torch.randn(D, 2 * D, device=dev) * 0.05; b2 = torch.zeros(D, device=dev)
g = torch.randn(M, D, device=dev)
for _ in range(3):
    z1 = ops._linear_fwd(x, w1, b1, True)        # fwd  [M,300]x[300,600]
    z2 = ops._linear_fwd(z1, w2, b2, False)      # fwd  [M,600]x[600,300]
    gw2, gb2 = ops._linear_bwd_w(g, z1)          # wgrad 300x600 over M
    gz1 = ops._linear_bwd_x(g, w2, z1)           # dgrad [M,300]x[300,600]
    gw1, gb1 = ops._linear_bwd_w(gz1, x)         # wgrad 600x300 over M
    ga = ops._linear_bwd_x(gz1, w1)              # dgrad [M,600]x[600,300]
torch.cuda.synchronize()
print("done")
