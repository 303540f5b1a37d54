"""One launch of each hand-written conv kernel at the config-2 shapes (for `ncu --set full`)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
PKG = importlib.import_module("cv_a-fan_b200")
conv = PKG.conv
dev = torch.device("cuda:0")
for (n, c, h) in ((128, 16, 32), (128, 32, 16), (128, 64, 8), (256, 32, 16), (256, 64, 8)):
    m = conv.Conv3x3(c, c, 1).to(dev)
    x = torch.randn(n, c, h, h, device=dev, requires_grad=True)
    dy = torch.randn(n, c, h, h, device=dev)
    for _ in range(2):
        y = m(x)
        dx, dw = torch.autograd.grad(y, (x, m.weight), dy)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev).zero_()        # L2-cold inputs, like profiles/ncu_kernels.py
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    y = m(x)
    dx, dw = torch.autograd.grad(y, (x, m.weight), dy)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
