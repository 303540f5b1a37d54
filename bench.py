#!/usr/bin/env python
"""bench.py -- A-FAN training throughput (img/s) + hand-written-kernel rooflines on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's sm_100a path
    python bench.py --impl reference [--gpus N] [--steps K] ...      # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...                # one rank per GPU, NCCL, weak scaling

Workload (BASELINE.json configs[1]): ResNet-56 (resnet_s [9,9,9]) on CIFAR-100-shaped synthetic data,
A-FAN with PGD-5 on the layer-13 feature map (128x16x32x32 per GPU), random start + L-inf projection,
dual-BN tail, batch 128 PER GPU (1024 over 8 GPUs), fp32.  One "step" = one full training iteration
(Classification/main_perturb.py:173-201): head fwd, 5 x (tail fwd + dgrad + fused PGD step), [adv; clean]
dual-BN tail fwd, backward, gradient all-reduce, fused SGD.

Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = the same step through the
public API with pinned HOST inputs (H2D of images/labels and D2H of the loss inside the timed region).
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(net="resnet_s [9,9,9] (ResNet-56)", num_blocks=(9, 9, 9), num_classes=100, batch_per_gpu=128,
                image=(3, 32, 32), perturb_idx=13, steps=5, gamma=0.5, eps=2.0, randinit=True, clip=True)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = sorted(float(r[0]) for r in rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][1])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any("Active" in r[2 + i] and "Not" not in r[2 + i] for r in rows)]
            out["samples"] = len(rows)
        except Exception as e:       # no nvidia-smi (CPU box) -> nulls
            out["error"] = str(e)[:80]
        finally:
            if self.path and os.path.exists(self.path):
                os.unlink(self.path)
        return out


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU port of the reference iteration (oracle/afan_ref_torch.py)
# ---------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, batch=None):
    from oracle import afan_ref_torch as ref_t
    w = WORKLOAD
    batch = batch or w["batch_per_gpu"]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(3)
    model = ref_t.CifarResNetRef(w["num_blocks"], w["num_classes"])
    model.train()
    opt, crit = ref_t.make_sgd(model), torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(batch, *w["image"], generator=g)
    y = torch.randint(0, w["num_classes"], (batch,), generator=g)
    kw = dict(steps=w["steps"], gamma=w["gamma"], eps=w["eps"], perturb_idx=w["perturb_idx"],
              randinit=w["randinit"], clip=w["clip"])
    for _ in range(warmup):
        ref_t.afan_train_iteration(model, opt, crit, x, y, **kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_t.afan_train_iteration(model, opt, crit, x, y, **kw)
    dt = time.perf_counter() - t0
    return {"value": batch * steps / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"{steps} full iterations (after {warmup} warm-up) of the same workload at batch {batch} "
                      f"on {cores} host threads, oracle/afan_ref_torch.py (plain-PyTorch port of main_perturb.py:173-201)",
            "ms_per_step": 1e3 * dt / steps}


def gpu_reference(dev, steps=5, warmup=3):
    """Context number: the reference iteration in plain PyTorch ON THE SAME B200 -- un-fused ATen PGD ops,
    nn.BatchNorm2d, two head passes, torch.optim.SGD, eager launches, cuDNN's fastest fp32 algorithms.  This is what the
    reference itself would run on this GPU; it is a baseline like `cpu_baseline`, never the product path."""
    from oracle import afan_ref_torch as ref_t
    w = WORKLOAD
    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = False
    try:
        torch.manual_seed(3)
        model = ref_t.CifarResNetRef(w["num_blocks"], w["num_classes"]).to(dev)
        model.train()
        opt, crit = ref_t.make_sgd(model), torch.nn.CrossEntropyLoss()
        g = torch.Generator().manual_seed(3)
        x = torch.rand(w["batch_per_gpu"], *w["image"], generator=g).to(dev)
        y = torch.randint(0, w["num_classes"], (w["batch_per_gpu"],), generator=g).to(dev)
        kw = dict(steps=w["steps"], gamma=w["gamma"], eps=w["eps"], perturb_idx=w["perturb_idx"], randinit=w["randinit"],
                  clip=w["clip"])
        for _ in range(warmup):
            ref_t.afan_train_iteration(model, opt, crit, x, y, **kw)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            ref_t.afan_train_iteration(model, opt, crit, x, y, **kw)
        e.record()
        e.synchronize()
        ms = s.elapsed_time(e) / steps
        return {"value": 1e3 * w["batch_per_gpu"] / ms, "unit": "img/s", "ms_per_step": ms,
                "what": "oracle/afan_ref_torch.py (plain-PyTorch port of main_perturb.py:173-201) on the same GPU: un-fused ATen "
                        "PGD, nn.BatchNorm2d, head forwarded twice, torch.optim.SGD, eager, cuDNN benchmark mode, fp32"}
    finally:
        torch.backends.cudnn.deterministic = det


def config_dict(n_gpus, conv_math="fp32", conv="afan"):
    w = WORKLOAD
    return {"workload": "BASELINE configs[1]: ResNet-56 CIFAR-100-shaped synthetic 32x32, A-FAN PGD-5 with dual BN",
            "global_batch": w["batch_per_gpu"] * n_gpus, "batch_per_gpu": w["batch_per_gpu"],
            "perturb_idx": w["perturb_idx"], "perturbed_feature": "128x16x32x32 fp32 per GPU",
            "pgd_steps": w["steps"], "gamma_255": w["gamma"], "eps_255": w["eps"], "randinit": w["randinit"],
            "clip": w["clip"], "parallelism": f"dp{n_gpus} (one process per GPU, NCCL)",
            "conv_math": "fp32 (TF32 off)" if conv_math == "fp32" else "tf32 tensor cores (cuDNN), fp32 accumulate",
            "conv3x3": {"afan": "hand-written sm_100a direct convolution (strict fp32 FFMA)" if conv_math == "fp32"
                        else "hand-written sm_100a mma.sync TF32 implicit GEMM (1 pass)",
                        "3xtf32": "hand-written sm_100a mma.sync 3xTF32 implicit GEMM", "cudnn": "cuDNN"}[conv],
            "deterministic": "cudnn.deterministic=True (main_perturb.py:315) + deterministic hand-written kernels: bitwise-reproducible steps",
            "l2": "no flush between steps: per-step working set (saved activations ~0.9 GB) exceeds the 126 MB L2; "
                  "kernel rooflines are measured separately with an L2 flush between launches"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "A-FAN train img/s", "value": r["value"], "unit": "img/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# kernel rooflines: the hand-written kernels at the workload's shapes, L2-cold
# ---------------------------------------------------------------------------------------------------
L2_BYTES = 126 * 1024 * 1024


def kernel_rooflines(pkg, dev, reps=10):
    """Per-launch device time of each hand-written kernel with an L2-cold working set.

    Method ("inputs larger than L2"): every kernel gets R independent tensor sets with R x bytes >= 4 x L2;
    one CUDA graph launches the kernel once per set back to back, the graph is replayed `reps` times between
    two CUDA events on the launching stream, and the per-launch time is total / (R x reps).  Each launch
    therefore finds none of its inputs in L2 and no CPU launch latency is in the number; inter-kernel gaps
    are (they are real cost in the training step as well).  `us_isolated` is the classic single launch
    between two events after a 256 MB L2-flush memset (includes event/launch ramp, shown for context)."""
    ops = pkg.ops
    L = pkg._lib.lib()
    st = pkg._lib.stream
    peak, peak_src = peaks()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    side = torch.cuda.Stream(device=dev)
    res = []

    # fp32 FFMA peak of this device at its current clocks (roofline denominator of the direct convolutions)
    probe_out = torch.empty(2 * 148 * 256 * 2, dtype=torch.float32, device=dev)
    fl = ctypes.c_double(0.0)
    for _ in range(2):
        pkg._lib.check(L.afan_ffma_probe(probe_out.data_ptr(), probe_out.numel(), 20000, ctypes.byref(fl), st()), "afan_ffma_probe")
    torch.cuda.synchronize()
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    pkg._lib.check(L.afan_ffma_probe(probe_out.data_ptr(), probe_out.numel(), 20000, ctypes.byref(fl), st()), "afan_ffma_probe")
    e0.record()
    e0.synchronize()
    ffma_peak = fl.value / (s0.elapsed_time(e0) * 1e-3) / 1e12

    def measure(name, bytes_per_launch, footprint, make_set, run, launches_per_iter, note, flops=None):
        R = max(2, min(64, -(-4 * L2_BYTES // footprint)))
        sets = [make_set() for _ in range(R)]
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for sset in sets:
                run(sset)                               # warm-up (lazy module load) outside capture
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for sset in sets:
                run(sset)
        for _ in range(3):
            graph.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(reps):
            graph.replay()
        e.record()
        e.synchronize()
        t = s.elapsed_time(e) * 1e-3 / (R * reps)
        iso = []
        for i in range(8):
            flush.zero_()
            s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s2.record()
            run(sets[i % R])
            e2.record()
            e2.synchronize()
            if i >= 3:
                iso.append(s2.elapsed_time(e2) * 1e-3)
        row = {"kernel": name, "family": name.split(" [")[0].split("+")[0].split(" clip")[0].split(" noclip")[0].strip(),
               "bytes": bytes_per_launch, "us": t * 1e6, "us_isolated": 1e6 * sum(iso) / len(iso),
               "rotating_sets": R, "launches_per_iter": launches_per_iter, "note": note}
        if flops is None:
            row.update({"bound": "hbm", "achieved": bytes_per_launch / t / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": bytes_per_launch / t / 1e9 / peak})
        else:
            row.update({"bound": "fp32_ffma", "flops": flops, "achieved": flops / t / 1e12, "peak": ffma_peak, "unit": "TFLOP/s",
                        "frac": flops / t / 1e12 / ffma_peak, "hbm_gbs": bytes_per_launch / t / 1e9})
        res.append(row)
        del graph, sets

    w = WORKLOAD
    n = w["batch_per_gpu"]
    gm, ep = w["gamma"] / 255, w["eps"] / 255
    for tag, shape in (("cfg2 128x16x32x32", (n, 16, 32, 32)), ("cfg4 8x1024x38x63", (8, 1024, 38, 63))):
        E = 1
        for d in shape:
            E *= d
        in_step = tag.startswith("cfg2")

        def mk():
            x = torch.relu(1.5 * torch.randn(shape, device=dev, generator=g))
            return dict(x=x, g=1e-3 * torch.randn(shape, device=dev, generator=g), xa=x.clone(), d=torch.empty_like(x),
                        nrm=torch.zeros(2, shape[0], device=dev), ws=ops.norms_workspace(shape[0], dev))
        measure(f"pgd_linf_step clip [{tag}]", 16 * E, 12 * E, mk,
                lambda t: ops.pgd_linf_step_(t["g"], t["x"], t["xa"], gm, ep, True), w["steps"] - 1 if in_step else 0,
                "read g,x_adv,x + write x_adv = 16 B/elem")
        measure(f"pgd_linf_step clip+norms [{tag}]", 16 * E, 12 * E, mk,
                lambda t: ops.pgd_linf_step_(t["g"], t["x"], t["xa"], gm, ep, True, norms_out=t["nrm"], workspace=t["ws"]),
                1 if in_step else 0, "16 B/elem, per-sample L2/Linf norms fused (last PGD step of the trainer)")
        measure(f"pgd_linf_step clip+delta+norms [{tag}]", 20 * E, 16 * E, mk,
                lambda t: ops.pgd_linf_step_(t["g"], t["x"], t["xa"], gm, ep, True, delta_out=t["d"], norms_out=t["nrm"],
                                             workspace=t["ws"]), 0, "+ write delta = 20 B/elem")
        measure(f"pgd_linf_step noclip [{tag}]", 12 * E, 8 * E, mk,
                lambda t: ops.pgd_linf_step_(t["g"], None, t["xa"], gm, ep, False), 0, "shipped recipe (no clip): 12 B/elem")
        measure(f"pgd_init philox [{tag}]", 8 * E, 8 * E, mk, lambda t: ops.pgd_init(t["x"], ep, seed=1, out=t["xa"]),
                1 if in_step else 0, "read x + write x_adv = 8 B/elem, noise generated in registers")
    for tag, (G, N, C, H, W), lpi in (("G2 128x32x16x16", (2, n, 32, 16, 16), 18), ("G2 128x64x8x8", (2, n, 64, 8, 8), 18),
                                      ("G1 128x32x16x16", (1, n, 32, 16, 16), 18 * w["steps"]),
                                      ("G1 128x64x8x8", (1, n, 64, 8, 8), 18 * w["steps"]),
                                      ("G1 128x16x32x32", (1, n, 16, 32, 32), 19),
                                      ("G2 256x64x56x56", (2, 256, 64, 56, 56), 0)):
        E = G * N * C * H * W
        nb = L.afan_bn_workspace_bytes(G, C)

        def mk():
            return dict(x=torch.randn(G * N, C, H, W, device=dev, generator=g), dy=torch.randn(G * N, C, H, W, device=dev, generator=g),
                        y=torch.empty(G * N, C, H, W, device=dev), dx=torch.empty(G * N, C, H, W, device=dev),
                        w=torch.ones(C, device=dev), b=torch.zeros(C, device=dev), rm=torch.zeros(C, device=dev),
                        rv=torch.ones(C, device=dev), sm=torch.empty(G, C, device=dev), si=torch.empty(G, C, device=dev),
                        dw=torch.empty(C, device=dev), db=torch.empty(C, device=dev), ws=ops.bn_workspace(G, C, dev))

        def fwd(t):
            pkg._lib.check(L.afan_bn_fwd_f32(t["x"].data_ptr(), None, t["w"].data_ptr(), t["b"].data_ptr(), t["rm"].data_ptr(),
                                             t["rv"].data_ptr(), t["y"].data_ptr(), t["sm"].data_ptr(), t["si"].data_ptr(),
                                             t["ws"].data_ptr(), nb, G, N, C, H * W, 1e-5, 0.1, 1, 1, st()), "afan_bn_fwd_f32")

        def bwd(t):
            pkg._lib.check(L.afan_bn_bwd_f32(t["dy"].data_ptr(), t["x"].data_ptr(), t["y"].data_ptr(), t["w"].data_ptr(),
                                             t["sm"].data_ptr(), t["si"].data_ptr(), t["dx"].data_ptr(), None, t["dw"].data_ptr(),
                                             t["db"].data_ptr(), t["ws"].data_ptr(), nb, G, N, C, H * W, 1, st()), "afan_bn_bwd_f32")

        def mk_bwd():
            t = mk()
            fwd(t)
            return t
        measure(f"dual_bn fwd+relu [{tag}]", 8 * E, 8 * E, mk, fwd, lpi,
                "read x + write y = 8 B/elem (16 B per clean+adv pair)")
        measure(f"dual_bn bwd+relu [{tag}]", 16 * E, 16 * E, mk_bwd, bwd, lpi,
                "read dy,x,y + write dx = 16 B/elem (12 without the ReLU mask)")
    # 3x3 convolutions of the step (ResNet-56, perturb_idx 13): stage-1 head at batch n, tail stages 2/3 at batch n in the
    # ascent (forward + dgrad per PGD step) and at 2n in the final [adv; clean] pass
    steps = w["steps"]
    for (N, C, H), lpi_conv, lpi_wgrad in (((n, 16, 32), 36, 18), ((n, 32, 16), 34 * steps, 0), ((n, 64, 8), 34 * steps, 0),
                                           ((2 * n, 32, 16), 34, 17), ((2 * n, 64, 8), 34, 17)):
        E = N * C * H * H
        fl_conv = 2.0 * E * C * 9

        def mk():
            m = pkg.conv.Conv3x3(C, C, 1).to(dev)
            wf, wd = m.packed()
            return dict(x=torch.randn(N, C, H, H, device=dev, generator=g), dy=torch.randn(N, C, H, H, device=dev, generator=g),
                        wf=wf, m=m, ws=m.wgrad_workspace())
        math = {"afan": "fp32", "tf32": "tf32", "3xtf32": "3xtf32"}.get(pkg.conv.MODE, "fp32")
        measure(f"conv3x3 fwd/dgrad [{N}x{C}x{H}x{H}]", 8 * E + 4 * 9 * C * C, 8 * E, mk,
                lambda t: ops.conv3x3(t["x"], t["wf"], math=math), lpi_conv,
                "3x3 s1 p1 conv, forward and (other weight packing) input gradient: 18*C FLOP per output element; "
                "read x + write y = 8 B/elem", flops=fl_conv)
        measure(f"conv3x3 wgrad [{N}x{C}x{H}x{H}]", 8 * E + 4 * 9 * C * C, 8 * E, mk,
                lambda t: ops.conv3x3_wgrad(t["x"], t["dy"], t["ws"]), lpi_wgrad,
                "weight gradient (partials kernel + fixed-order fold kernel, both in the time)", flops=fl_conv)
    # stride-2 stage transitions (first block of stages 2 and 3): forward + dgrad per PGD step at batch n, once at 2n
    for (N, CIN, HIN), lpi in (((n, 16, 32), steps), ((n, 32, 16), steps), ((2 * n, 16, 32), 1), ((2 * n, 32, 16), 1)):
        HO = HIN // 2
        fl = 2.0 * N * HO * HO * (2 * CIN) * CIN * 9
        nbytes = 4 * N * CIN * HIN * HIN + 4 * N * 2 * CIN * HO * HO

        def mk():
            m = pkg.conv.Conv3x3(CIN, 2 * CIN, 2).to(dev)
            wf, wd = m.packed()
            return dict(x=torch.randn(N, CIN, HIN, HIN, device=dev, generator=g), dy=torch.randn(N, 2 * CIN, HO, HO, device=dev, generator=g),
                        wf=wf, wd=wd, m=m)
        measure(f"conv3x3s2 fwd [{N}x{CIN}x{HIN}x{HIN}]", nbytes, nbytes, mk, lambda t: ops.conv3x3s2(t["x"], t["wf"]), lpi,
                "stride-2 transition C -> 2C: 36*C FLOP per output element", flops=fl)
        measure(f"conv3x3s2 dgrad [{N}x{CIN}x{HIN}x{HIN}]", nbytes, nbytes, mk, lambda t: ops.conv3x3s2(t["dy"], t["wd"], dgrad=True), lpi,
                "input gradient of the stride-2 transition", flops=fl)
    for tag, shape in (("cfg5 4x2048x33x33", (4, 2048, 33, 33)), ("cfg4 8x1024x38x63", (8, 1024, 38, 63)),
                       ("4x256x128x128", (4, 256, 128, 128))):
        E = 1
        for d in shape:
            E *= d

        def mk():
            cl = torch.relu(torch.randn(shape, device=dev, generator=g))
            return dict(cl=cl, ad=cl + 0.01 * torch.randn(shape, device=dev, generator=g), out=torch.empty_like(cl))
        measure(f"mix_feature [{tag}]", 12 * E, 12 * E, mk, lambda t: ops.mix_feature(t["cl"], t["ad"], out=t["out"]), 0,
                "read clean, adv + write out = 12 B/elem (Seg/Det normalisation; not in the Classification step)")
    return res, peak_src, ffma_peak


def exchange_microbench(pkg, dev, mailbox, pg, calls=100):
    """N>1 only, every rank: per-call device time of the dual-BN forward at a tail shape (G1 128x32x16x16) with the
    fused NVLink exchange, with the NCCL split form, and without any exchange (graph of `calls` back-to-back launches)."""
    ops = pkg.ops
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(128, 32, 16, 16, device=dev, generator=g)
    w, b = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    rm, rv = torch.zeros(32, device=dev), torch.ones(32, device=dev)
    ws = ops.bn_workspace(1, 32, dev)
    side = torch.cuda.Stream(device=dev)
    out = {}
    for name, kw in (("local_no_exchange", {}), ("fused_nvlink_p2p", dict(mailbox=mailbox)), ("nccl_split", dict(process_group=pg))):
        if name == "fused_nvlink_p2p" and mailbox is None:
            continue
        run = lambda: ops.bn_fwd(x, None, w, b, rm, rv, ws, groups=1, relu=True, **kw)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.distributed.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(calls):
                run()
        graph.replay()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(5):
            graph.replay()
        e0.record()
        e0.synchronize()
        out[name] = s0.elapsed_time(e0) * 1e3 / (5 * calls)
        torch.distributed.barrier()
        del graph
    return out


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="afan_b200", choices=["afan_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    ap.add_argument("--no-sync-bn", action="store_true", help="per-replica BN statistics (the reference's DataParallel behaviour)")
    ap.add_argument("--bn-exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU dual-BN statistics: fused NVLink peer-memory exchange inside the kernel, or NCCL all-reduce")
    ap.add_argument("--conv-math", default="fp32", choices=["fp32", "tf32"],
                    help="cuDNN/cuBLAS math of the (library) convolutions / fc: strict fp32 (headline) or TF32 tensor cores")
    ap.add_argument("--conv", default="afan", choices=["afan", "cudnn", "3xtf32"],
                    help="3x3 tail convolutions: hand-written sm_100a kernels (default; strict-fp32 FFMA, or the TF32 "
                         "tensor-core twin under --conv-math tf32), the cuDNN library path, or the 3xTF32 split kernels")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-variants", action="store_true", help="do not time the TF32 / cuDNN convolution variants of the step")
    ap.add_argument("--skip-rooflines", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager iteration between cudaProfilerStart/Stop (for `ncu --profile-from-start off`) and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the afan_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
        pg = torch.distributed.group.WORLD
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    torch.backends.cudnn.allow_tf32 = args.conv_math == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = args.conv_math == "tf32"
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.deterministic = os.environ.get("AFAN_BENCH_DET", "1") != "0"   # the reference's setting (main_perturb.py:315): bitwise-reproducible steps

    pkg = importlib.import_module("cv_a-fan_b200")
    pkg.conv.MODE = "tf32" if (args.conv == "afan" and args.conv_math == "tf32") else args.conv
    w = WORKLOAD
    torch.manual_seed(3)                                     # identical weights on every rank
    model = pkg.resnet_s.ResNet(num_blocks=w["num_blocks"], num_classes=w["num_classes"]).to(dev)
    trainer = pkg.trainer.AfanTrainer(model, perturb_idx=w["perturb_idx"], steps=w["steps"], gamma=w["gamma"],
                                      eps=w["eps"], randinit=w["randinit"], clip=w["clip"], rng="philox", seed=3 + rank,
                                      process_group=pg, sync_bn=not args.no_sync_bn, use_cuda_graph=not args.no_graph,
                                      bn_exchange=args.bn_exchange)
    n = w["batch_per_gpu"]
    g = torch.Generator().manual_seed(3 + rank)              # per-rank data
    host_x = [torch.rand(n, *w["image"], generator=g).pin_memory() for _ in range(4)]
    host_y = [torch.randint(0, w["num_classes"], (n,), generator=g).pin_memory() for _ in range(4)]
    dev_x, dev_y = [t.to(dev) for t in host_x], [t.to(dev) for t in host_y]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed_run(step_fn):
        for i in range(args.warmup):
            step_fn(i)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            s.record()
            for i in range(args.steps):
                step_fn(i)
            e.record()
            barrier()
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / 1e3, clk.summary()

    if args.profile_step:
        eager = pkg.trainer.AfanTrainer(model, perturb_idx=w["perturb_idx"], steps=w["steps"], gamma=w["gamma"],
                                        eps=w["eps"], randinit=w["randinit"], clip=w["clip"], rng="philox", seed=3,
                                        use_cuda_graph=False)
        for i in range(3):
            eager.step(dev_x[i], dev_y[i])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eager.step(dev_x[3], dev_y[3])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_step": True, "afan_kernels_per_step": eager.kernel_launches_per_iter}))
        return

    # (1) device-resident inputs
    trainer.step(dev_x[0], dev_y[0])                         # builds arena, captures the graph
    out = {}

    def step_dev(i):
        out["r"] = trainer.step(dev_x[i % 4], dev_y[i % 4])
    sec, clocks = timed_run(step_dev)
    loss_dev = float(out["r"]["loss"])

    # (2) end to end: pinned host inputs -> H2D -> step -> D2H loss
    host_loss = torch.zeros(1).pin_memory()
    dx, dy = torch.empty_like(dev_x[0]), torch.empty_like(dev_y[0])

    def step_e2e(i):
        dx.copy_(host_x[i % 4], non_blocking=True)
        dy.copy_(host_y[i % 4], non_blocking=True)
        r = trainer.step(dx, dy)
        host_loss.copy_(r["loss"].reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the user reads the loss every step (main_perturb.py:208)
    sec_e2e, _ = timed_run(step_e2e)

    per_iter = trainer.kernel_launches_per_iter      # afan kernels per iteration, counted at capture/trace time

    global_batch = n * world
    line = {"metric": "A-FAN train img/s", "value": global_batch * args.steps / sec, "unit": "img/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world, args.conv_math, args.conv), "clocks": clocks,
            "e2e": {"value": global_batch * args.steps / sec_e2e, "unit": "img/s",
                    "h2d_bytes_per_step": (host_x[0].numel() * 4 + host_y[0].numel() * 8) * world,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": 1e3 * sec_e2e / args.steps},
            "gpu_launches": per_iter * args.steps, "afan_kernels_per_step": per_iter,
            "cuda_graph": not args.no_graph, "sync_bn": (not args.no_sync_bn) and world > 1,
            "bn_exchange": args.bn_exchange if ((not args.no_sync_bn) and world > 1) else None, "final_loss": loss_dev}

    # other convolution paths of the SAME step, for context only (never the headline): single GPU, device-resident inputs
    if world == 1 and not args.skip_variants and args.conv == "afan" and args.conv_math == "fp32" and not args.no_graph:
        variants = {}
        det = torch.backends.cudnn.deterministic
        for vname, mode, tf32 in (("conv_tf32_tensor_core_kernels", "tf32", True), ("conv_cudnn_fp32_nondeterministic", "cudnn", False)):
            pkg.conv.MODE = mode
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            # the library path is timed with its fastest (non-deterministic) algorithms: under cudnn.deterministic the same
            # step takes 25.4 ms
            torch.backends.cudnn.deterministic = det and mode != "cudnn"
            torch.manual_seed(3)
            vm = pkg.resnet_s.ResNet(num_blocks=w["num_blocks"], num_classes=w["num_classes"]).to(dev)
            vt = pkg.trainer.AfanTrainer(vm, perturb_idx=w["perturb_idx"], steps=w["steps"], gamma=w["gamma"], eps=w["eps"],
                                         randinit=w["randinit"], clip=w["clip"], rng="philox", seed=3, use_cuda_graph=True)
            vt.step(dev_x[0], dev_y[0])
            vsec, _ = timed_run(lambda i: vt.step(dev_x[i % 4], dev_y[i % 4]))
            variants[vname] = {"value": n * args.steps / vsec, "unit": "img/s", "ms_per_step": 1e3 * vsec / args.steps}
            vt.close()
            del vm, vt
        pkg.conv.MODE = "afan"
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.deterministic = det
        line["variants"] = variants

    if rank == 0 and not args.skip_rooflines:
        ks, peak_src, ffma_peak = kernel_rooflines(pkg, dev)
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tmap = json.load(f)
        except Exception:
            tmap = {}
        for k in ks:
            k["traffic"] = tmap.get(k["kernel"])
        # Roofline per kernel FAMILY (all in-step shapes of a kernel, weighted by their launches in one step):
        # achieved = algorithmic bytes (HBM-bound families) or FLOPs (the FFMA-bound convolutions) per step / device time
        # per step.  `roofline` = the family with the largest share of the step; `roofline_hbm` = the largest HBM-bound one.
        step_us = 1e3 * sec / args.steps * 1e3
        fam = {}
        for k in ks:
            if k["launches_per_iter"] > 0:
                f = fam.setdefault(k["family"], {"bound": k["bound"], "peak": k["peak"], "unit": k["unit"], "work": 0.0, "bytes": 0.0,
                                                 "us": 0.0, "launches": 0, "traffic": 0.0, "traffic_ok": True, "shapes": []})
                f["work"] += (k["flops"] if k["bound"] == "fp32_ffma" else k["bytes"]) * k["launches_per_iter"]
                f["bytes"] += k["bytes"] * k["launches_per_iter"]
                f["us"] += k["us"] * k["launches_per_iter"]
                f["launches"] += k["launches_per_iter"]
                f["shapes"].append(k["kernel"])
                if k["traffic"] is None:
                    f["traffic_ok"] = False
                else:
                    f["traffic"] += k["traffic"] * k["launches_per_iter"]

        def fam_line(name, f):
            scale = 1e12 if f["bound"] == "fp32_ffma" else 1e9
            achieved = f["work"] / (f["us"] * 1e-6) / scale
            d = {"bound": f["bound"], "kernel": name, "achieved": achieved, "peak": f["peak"], "unit": f["unit"],
                 "frac": achieved / f["peak"], "traffic": f["traffic"] / f["launches"] if f["traffic_ok"] else None,
                 "algorithmic_bytes": f["bytes"] / f["launches"], "launches_per_step": f["launches"],
                 "avg_us_per_launch": f["us"] / f["launches"], "share_of_step": f["us"] / step_us, "shapes": f["shapes"]}
            if f["bound"] == "fp32_ffma":
                d["algorithmic_flops"] = f["work"] / f["launches"]
                d["peak_source"] = ("fp32 FFMA rate measured live by afan_ffma_probe (8x8 outer-product loop, 2 CTAs x 256 "
                                    "threads per SM) on this device at its current clocks; nominal 148 SMs x 128 lanes x 2 x "
                                    "1.965 GHz = 74.4 TFLOP/s")
            else:
                d["peak_source"] = peak_src
            return d
        timing = ("per shape: CUDA events around a graph of back-to-back launches over rotating tensor sets (4x L2, every launch "
                  "L2-cold), weighted by the launches of that shape in one step; traffic / algorithmic_bytes are per-launch averages")
        dom_name, dom = max(fam.items(), key=lambda kv: kv[1]["us"])
        line["roofline"] = fam_line(dom_name, dom)
        line["roofline"]["timing"] = timing
        line["roofline"]["families"] = {n: {k2: v for k2, v in fam_line(n, f).items()
                                            if k2 in ("bound", "achieved", "peak", "unit", "frac", "share_of_step", "launches_per_step")}
                                        for n, f in fam.items()}
        hbm = {n: f for n, f in fam.items() if f["bound"] == "hbm"}
        if hbm and dom["bound"] != "hbm":
            hn, hf = max(hbm.items(), key=lambda kv: kv[1]["us"])
            line["roofline_hbm"] = fam_line(hn, hf)
        line["ffma_peak_tflops_measured"] = ffma_peak
        line["kernels"] = ks
    if world > 1:
        torch.distributed.barrier()
        if not args.no_sync_bn:
            line["bn_fwd_us_per_call_G1_128x32x16x16"] = exchange_microbench(pkg, dev, trainer.mailbox, pg)
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        try:
            line["reference_on_gpu"] = gpu_reference(dev)
        except Exception as exc:                     # context number only: never fail the bench line over it
            line["reference_on_gpu"] = {"error": repr(exc)[:200]}
        cb = cpu_reference(steps=4, warmup=1)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        trainer.close()                      # a live graph with captured NCCL ops blocks communicator teardown
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
