import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("cv_a-fan_b200")
dev = torch.device("cuda:0")
for (n, cin, hin) in ((128, 32, 16), (128, 16, 32)):
    m = pkg.conv.Conv3x3(cin, 2 * cin, 2).to(dev)
    x = torch.randn(n, cin, hin, hin, device=dev)
    wf, wd = m.packed()
    for _ in range(3):
        y = pkg.ops.conv3x3s2(x, wf)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    y = pkg.ops.conv3x3s2(x, wf)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
