import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cv_a-fan_b200"); ops = pkg.ops
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(3)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for G, N, C, H, W in ((1, 128, 16, 32, 32), (1, 128, 32, 16, 16), (2, 128, 64, 8, 8)):
    x = torch.randn(G * N, C, H, W, device=dev, generator=g); dy = torch.randn(G * N, C, H, W, device=dev, generator=g)
    w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev); rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ws = ops.bn_workspace(G, C, dev)
    for _ in range(2):
        flush.zero_(); y, sm, si = ops.bn_fwd(x, None, w, b, rm, rv, ws, groups=G, relu=True); torch.cuda.synchronize()
        flush.zero_(); ops.bn_bwd(dy, x, y, w, sm, si, ws, groups=G, relu=True); torch.cuda.synchronize()
print("done")
