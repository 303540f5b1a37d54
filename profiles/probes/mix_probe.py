"""CUDA-event timing (L2-cold rotating sets) of mix_feature at the config shapes; AFAN_MIX_DB selects the double-buffered forms."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = importlib.import_module("cv_a-fan_b200")
dev = torch.device("cuda:0")
PEAK = 6547.8
for shape in ((4, 2048, 33, 33), (8, 1024, 38, 63), (4, 256, 128, 128), (4, 256, 129, 129), (2, 304, 129, 129)):
    n = 1
    for d in shape:
        n *= d
    sets = min(24, max(2, int(4 * 126e6 / (n * 4 * 3)) + 1))
    cl = [torch.relu(torch.randn(shape, device=dev)) for _ in range(sets)]
    ad = [c + 0.01 * torch.randn(shape, device=dev) for c in cl]
    out = torch.empty(shape, device=dev)
    # replayed from ONE CUDA graph: a Python call (+ the scratch allocation of the row-streaming form) costs more than the kernels
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            PKG.ops.mix_feature(cl[i % sets], ad[i % sets], out=out)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    reps = 48
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for i in range(reps):
            PKG.ops.mix_feature(cl[i % sets], ad[i % sets], out=out)
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record(); b.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    del graph
    c0, a0 = cl[0].double(), ad[0].double()
    want = (c0 - c0.mean(1, keepdim=True)) / (c0.var(1, keepdim=True) + 1e-5).sqrt() * (a0.var(1, keepdim=True) + 1e-5).sqrt() + a0.mean(1, keepdim=True)
    err = float((PKG.ops.mix_feature(cl[0], ad[0]).double() - want).abs().max())
    print(shape, "us", round(us, 1), "frac", round(n * 12 / us / 1e3 / PEAK, 3), "max_abs_err_vs_fp64", f"{err:.2e}", flush=True)
    del cl, ad
