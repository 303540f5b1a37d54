"""One launch of the tensor-core conv kernel (1 and 3 passes) at the config-2 shapes (for `ncu --set full`)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
PKG = importlib.import_module("cv_a-fan_b200")
_lib = PKG._lib
L = _lib.lib()
dev = torch.device("cuda:0")
for (n, c, h) in ((128, 32, 16), (128, 64, 8), (128, 16, 32)):
    x = torch.randn(n, c, h, h, device=dev)
    w = torch.randn(c, c, 3, 3, device=dev)
    wtf, wtd = torch.empty(2 * c * 9 * c, device=dev), torch.empty(2 * c * 9 * c, device=dev)
    desc = torch.tensor([[w.data_ptr(), wtf.data_ptr(), wtd.data_ptr(), c]], dtype=torch.int64, device=dev)
    y = torch.empty_like(x)
    for passes in (1, 3):
        L.afan_conv3x3_pack_tc_f32(desc.data_ptr(), 1, c, passes, _lib.stream())
        for _ in range(2):
            L.afan_conv3x3_tc_f32(x.data_ptr(), wtf.data_ptr(), y.data_ptr(), None, n, c, h, passes, 0, _lib.stream())
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        L.afan_conv3x3_tc_f32(x.data_ptr(), wtf.data_ptr(), y.data_ptr(), None, n, c, h, passes, 0, _lib.stream())
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
