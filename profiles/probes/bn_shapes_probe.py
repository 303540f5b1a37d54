"""CUDA-event timing (L2-cold rotating sets) of dual-BN fwd / bwd at the odd-H*W shapes of config 5."""
import importlib, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = importlib.import_module("cv_a-fan_b200")
ops = PKG.ops
dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6547.8) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6547.8
SHAPES = ((2, 4, 256, 129, 129), (2, 4, 256, 33, 33), (1, 4, 2048, 33, 33), (1, 8, 512, 75, 125), (2, 256, 64, 56, 56))
if os.environ.get("AFAN_PROBE_SHAPE"):
    SHAPES = (SHAPES[int(os.environ["AFAN_PROBE_SHAPE"])],)
for (G, N, C, H, W) in SHAPES:
    nb = G * N
    elems = nb * C * H * W
    sets = max(2, int(4 * 126e6 * 2 / (elems * 4 * 2)) + 1)
    sets = min(sets, 24)
    xs = [torch.randn(nb, C, H, W, device=dev) for _ in range(sets)]
    dys = [torch.randn(nb, C, H, W, device=dev) for _ in range(sets)]
    w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    ws = ops.bn_workspace(G, C, dev)
    def fwd(i):
        return ops.bn_fwd(xs[i % sets], None, w, b, rm, rv, ws, groups=G, relu=True)
    y, sm, si = fwd(0)
    def bwd(i):
        return ops.bn_bwd(dys[i % sets], xs[i % sets], y, w, sm, si, ws, groups=G, relu=True)
    res = {}
    for name, fn, bytes_per in (("fwd", fwd, 8), ("bwd", bwd, 16)):
        # the launches are replayed from ONE CUDA graph: a Python call costs ~20 us, more than the small shapes' kernels
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                fn(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        reps = 40
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for i in range(reps):
                fn(i)
        graph.replay()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        e.record(); e.synchronize()
        us = a.elapsed_time(e) * 1e3 / reps
        res[name] = (round(us, 1), round(elems * bytes_per / us / 1e3 / PEAK, 3))
        del graph
    print((G, N, C, H, W), res, flush=True)
    del xs, dys
