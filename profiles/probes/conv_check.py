"""GPU probe: hand-written conv3x3 (fwd / dgrad) vs cuDNN fp32 -- correctness against an fp64 reference and
back-to-back graph timing.  Usage: python profiles/probes/conv_check.py"""
import importlib, json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("cv_a-fan_b200")
from importlib import import_module
_lib = import_module("cv_a-fan_b200._lib")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")


def pack(w):
    C = w.shape[0]
    wf, wd = torch.empty(C * 9 * C, device=dev), torch.empty(C * 9 * C, device=dev)
    desc = torch.tensor([[w.data_ptr(), wf.data_ptr(), wd.data_ptr(), C]], dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().afan_conv3x3_pack_f32(desc.data_ptr(), 1, C, _lib.stream()), "pack")
    return wf, wd


def conv(x, wp, y, variant=0):
    n, c, h, w = x.shape
    _lib.check(_lib.lib().afan_conv3x3_f32(x.data_ptr(), wp.data_ptr(), y.data_ptr(), None, n, c, h, variant, _lib.stream()), "conv")
    return y


def graph_time(fn, reps=20, outer=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(outer):
        g.replay()
    b.record(); b.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * outer)


res = {}
for (n, c, h) in ((128, 16, 32), (128, 32, 16), (128, 64, 8), (256, 32, 16), (256, 64, 8), (4, 16, 16), (3, 64, 32), (5, 32, 8)):
    torch.manual_seed(0)
    x = torch.randn(n, c, h, h, device=dev)
    w = torch.randn(c, c, 3, 3, device=dev) * (2.0 / (9 * c)) ** 0.5
    dy = torch.randn(n, c, h, h, device=dev)
    wf, wd = pack(w)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), padding=1)
    cud = F.conv2d(x, w, padding=1)
    cud_dx = torch.nn.grad.conv2d_input(x.shape, w, dy, padding=1)
    y = torch.empty_like(x)
    key = f"{n}x{c}x{h}"
    res[key] = {"cudnn_fwd_err": (cud - ref).abs().max().item(), "cudnn_dgrad_err": (cud_dx - ref_dx).abs().max().item()}
    nvar = 4 if (n >= 128) else 1
    for v in range(nvar):
        conv(x, wf, y, v)
        e1 = (y - ref).abs().max().item()
        dx = torch.empty_like(x)
        conv(dy, wd, dx, v)
        e2 = (dx - ref_dx).abs().max().item()
        res[key][f"v{v}_fwd_err"] = e1
        res[key][f"v{v}_dgrad_err"] = e2
        if n >= 128:
            res[key][f"v{v}_us"] = round(graph_time(lambda: conv(x, wf, y, v)), 2)
    L = _lib.lib()
    wtf, wtd = torch.empty(2 * c * 9 * c, device=dev), torch.empty(2 * c * 9 * c, device=dev)
    desc = torch.tensor([[w.data_ptr(), wtf.data_ptr(), wtd.data_ptr(), c]], dtype=torch.int64, device=dev)
    def tc(inp, wp, out, passes, variant):
        _lib.check(L.afan_conv3x3_tc_f32(inp.data_ptr(), wp.data_ptr(), out.data_ptr(), None, n, c, h, passes, variant, _lib.stream()), "tc")
    for passes in (3, 1):
        _lib.check(L.afan_conv3x3_pack_tc_f32(desc.data_ptr(), 1, c, passes, _lib.stream()), "pack_tc")
        for v in range(2 if n >= 128 else 1):
            yt, dxt = torch.empty_like(x), torch.empty_like(x)
            tc(x, wtf, yt, passes, v); tc(dy, wtd, dxt, passes, v)
            res[key][f"tc{passes}_v{v}_fwd_err"] = (yt - ref).abs().max().item()
            res[key][f"tc{passes}_v{v}_dgrad_err"] = (dxt - ref_dx).abs().max().item()
            if n >= 128:
                res[key][f"tc{passes}_v{v}_us"] = round(graph_time(lambda: tc(x, wtf, yt, passes, v)), 2)
    wsb = L.afan_conv3x3_wgrad_workspace_bytes(c)
    ws = torch.empty(wsb // 4, device=dev)
    dw = torch.empty_like(w)
    def wgrad():
        _lib.check(L.afan_conv3x3_wgrad_f32(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), ws.data_ptr(), wsb, n, c, h, 0, _lib.stream()), "wgrad")
    wgrad()
    ref_dw = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), padding=1)
    cud_dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=1)
    res[key]["wgrad_relerr"] = ((dw - ref_dw).abs().max() / ref_dw.abs().max()).item()
    res[key]["cudnn_wgrad_relerr"] = ((cud_dw - ref_dw).abs().max() / ref_dw.abs().max()).item()
    if n >= 128:
        res[key]["wgrad_us"] = round(graph_time(wgrad), 2)
        res[key]["cudnn_wgrad_us"] = round(graph_time(lambda: torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=1)), 2)
    if n >= 128:
        res[key]["cudnn_fwd_us"] = round(graph_time(lambda: F.conv2d(x, w, padding=1)), 2)
        res[key]["cudnn_dgrad_us"] = round(graph_time(lambda: torch.nn.grad.conv2d_input(x.shape, w, dy, padding=1)), 2)
    print(key, json.dumps(res[key]), flush=True)

# stride-2 transitions: hand-written vs cuDNN
for (n, cin, hin) in ((128, 16, 32), (128, 32, 16), (256, 16, 32), (256, 32, 16)):
    m = pkg.conv.Conv3x3(cin, 2 * cin, 2).to(dev)
    x = torch.randn(n, cin, hin, hin, device=dev)
    dy = torch.randn(n, 2 * cin, hin // 2, hin // 2, device=dev)
    wf, wd = m.packed()
    ops = pkg.ops
    r = {"s2_fwd_us": round(graph_time(lambda: ops.conv3x3s2(x, wf)), 2),
         "s2_dgrad_us": round(graph_time(lambda: ops.conv3x3s2(dy, wd, dgrad=True)), 2),
         "cudnn_fwd_us": round(graph_time(lambda: F.conv2d(x, m.weight, stride=2, padding=1)), 2),
         "cudnn_dgrad_us": round(graph_time(lambda: torch.nn.grad.conv2d_input(x.shape, m.weight, dy, stride=2, padding=1)), 2)}
    print(f"s2 {n}x{cin}x{hin}", json.dumps(r), flush=True)
