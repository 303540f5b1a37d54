"""Timing probe: tcgen05 (UMMA) 3xTF32 convolution vs the strict-fp32 FFMA kernel at the tail shapes, L2-cold
(rotating tensor sets replayed from one CUDA graph, CUDA events).  python profiles/probes/umma_probe.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("cv_a-fan_b200")
ops, conv = pkg.ops, pkg.conv
dev = torch.device("cuda:0")


def time_graph(run, sets, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets:
            run(s)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for s in sets:
            run(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * len(sets))


SHAPES = [(128, 32, 16), (128, 64, 8), (256, 32, 16), (256, 64, 8)]
if os.environ.get("AFAN_PROBE_SHAPE"):                      # e.g. "128,64,8": one shape only (ncu captures)
    SHAPES = [tuple(int(v) for v in os.environ["AFAN_PROBE_SHAPE"].split(","))]
for n, c, h in SHAPES:
    out = {}
    for mode, math in (("afan", "fp32"), ("tc3", "umma")):
        conv.MODE = mode
        R = max(2, (4 * 126 * 2 ** 20) // (8 * n * c * h * h))
        sets = []
        for _ in range(min(R, 48)):
            m = conv.Conv3x3(c, c, 1).to(dev)
            wf, wd = m.packed()
            sets.append((torch.randn(n, c, h, h, device=dev), wf, m))
        out[mode] = time_graph(lambda s: ops.conv3x3(s[0], s[1], math=math), sets)
        if mode == "afan":
            ref = ops.conv3x3(sets[0][0], sets[0][1], math=math)
            x0, w0 = sets[0][0], sets[0][2].weight.detach().clone()
        else:
            m = conv.Conv3x3(c, c, 1).to(dev)
            with torch.no_grad():
                m.weight.copy_(w0)
            got = ops.conv3x3(x0, m.packed()[0], math=math)
            out["max_abs_diff_vs_ffma"] = float((got - ref).abs().max())
    print(f"conv3x3 {n}x{c}x{h}x{h}: ffma {out['afan']:.2f} us, tcgen05 3xtf32 {out['tc3']:.2f} us, "
          f"max |diff| {out['max_abs_diff_vs_ffma']:.2e}", flush=True)

# weight gradient: FFMA kernel (+ fold) vs tcgen05 kernel (+ fold)
for n, c, h in ([] if os.environ.get("AFAN_PROBE_SHAPE") else [(128, 32, 16), (256, 32, 16), (256, 64, 8)]):
    res = {}
    R = max(2, min(24, (4 * 126 * 2 ** 20) // (8 * n * c * h * h)))
    sets = [(torch.randn(n, c, h, h, device=dev), torch.randn(n, c, h, h, device=dev), ops.conv3x3_wgrad_workspace(c, dev),
             torch.zeros(c, c, 3, 3, device=dev)) for _ in range(R)]
    for math in ("fp32", "umma"):
        res[math] = time_graph(lambda s: ops.conv3x3_wgrad(s[0], s[1], s[2], accumulate_into=s[3], math=math), sets)
    a = ops.conv3x3_wgrad(sets[0][0], sets[0][1], sets[0][2], math="fp32")
    b = ops.conv3x3_wgrad(sets[0][0], sets[0][1], sets[0][2], math="umma")
    print(f"wgrad {n}x{c}x{h}x{h}: ffma {res['fp32']:.2f} us, tcgen05 3xtf32 {res['umma']:.2f} us, max |diff| {float((a - b).abs().max()):.2e} "
          f"(max |dw| {float(a.abs().max()):.1f})", flush=True)
