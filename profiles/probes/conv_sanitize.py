"""Small-shape launch of every convolution kernel family (for compute-sanitizer memcheck / racecheck)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("cv_a-fan_b200")
conv = pkg.conv
dev = torch.device("cuda:0")
torch.manual_seed(0)
for mode in ("afan", "tf32", "3xtf32"):
    conv.MODE = mode
    for (n, c, h) in ((3, 16, 32), (3, 32, 16), (3, 64, 8), (2, 64, 16)):
        m = conv.Conv3x3(c, c, 1).to(dev)
        x = torch.randn(n, c, h, h, device=dev, requires_grad=True)
        y, tap = m.forward_with_tap(x)
        dx, dw = torch.autograd.grad((y, tap), (x, m.weight), (torch.randn_like(y), torch.randn_like(x)))
conv.MODE = "afan"
for (n, c, h) in ((3, 16, 32), (3, 32, 16)):
    m = conv.Conv3x3(c, 2 * c, 2).to(dev)
    x = torch.randn(n, c, h, h, device=dev, requires_grad=True)
    y = m(x)
    dx, dw = torch.autograd.grad(y, (x, m.weight), torch.randn_like(y))
torch.cuda.synchronize()
print("done")
