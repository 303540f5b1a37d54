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
import threading
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
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (10 ms period, so a 0.3 s
    region still gets ~30 samples); `nvidia-smi -lms` in a subprocess when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path, self.thread, self.rows, self.max_mhz = index, None, None, None, [], None
        self._stop = threading.Event()

    def _nvml_loop(self, h, nv):
        bits = (nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap)
        while not self._stop.is_set():
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), [bool(r & b) for b in bits]))
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].strip().isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(h, nv), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        try:
            if self.thread is None:
                rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
                self.rows = [(float(r[0]), ["Active" in r[2 + i] and "Not" not in r[2 + i] for i in range(4)]) for r in rows]
                self.max_mhz = float(rows[0][1])
            sm = sorted(r[0] for r in self.rows)
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = self.max_mhz
            out["reasons"] = [n for i, n in enumerate(self.NAMES) if any(r[1][i] for r in self.rows)]
            out["samples"] = len(self.rows)
        except Exception as e:       # no NVML / nvidia-smi (CPU box) -> nulls
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
            "sample": f"{steps} full iterations (after {warmup} warm-up) of this workload, batch {batch}, {cores} threads, "
                      f"oracle/afan_ref_torch.py",
            "ms_per_step": 1e3 * dt / steps}


def gpu_reference(dev, steps=5, warmup=3, tf32=False):
    """Context number: the reference iteration in plain PyTorch ON THE SAME B200 -- un-fused ATen PGD ops,
    nn.BatchNorm2d, two head passes, torch.optim.SGD, eager launches, cuDNN's fastest fp32 algorithms.  This is what the
    reference itself would run on this GPU; it is a baseline like `cpu_baseline`, never the product path."""
    from oracle import afan_ref_torch as ref_t
    w = WORKLOAD
    det = torch.backends.cudnn.deterministic
    old_tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.deterministic = False
    torch.backends.cudnn.allow_tf32 = tf32           # stock PyTorch: cuDNN convolutions run in TF32 (allow_tf32 defaults to True)
    torch.backends.cuda.matmul.allow_tf32 = False    # stock PyTorch keeps fp32 matmul
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
                        "PGD, nn.BatchNorm2d, head forwarded twice, torch.optim.SGD, eager, cuDNN benchmark mode, "
                        + ("cuDNN TF32 convolutions (stock PyTorch defaults)" if tf32 else "strict fp32 convolutions")}
    finally:
        torch.backends.cudnn.deterministic = det
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32


def config_dict(n_gpus, conv_math="fp32", conv="afan", rng="philox"):
    """Short on purpose: the driver keeps only the tail of stdout, the whole final line must stay under ~1.4 KB."""
    w = WORKLOAD
    return {"workload": "configs[1]: ResNet-56 CIFAR-100-shaped, A-FAN PGD-5 idx13 rand+clip, dual BN",
            "global_batch": w["batch_per_gpu"] * n_gpus, "batch_per_gpu": w["batch_per_gpu"],
            "parallelism": f"dp{n_gpus}", "rng": rng,
            "conv": {"afan": "afan fp32 FFMA" if conv_math == "fp32" else "afan mma.sync tf32", "3xtf32": "afan mma.sync 3xtf32",
                     "tc3": "afan tcgen05 3xtf32", "cudnn": "cudnn " + conv_math}[conv],
            "deterministic": True, "l2": "step set > L2; kernel rooflines: rotating sets > 4x L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "A-FAN train img/s", "value": r["value"], "unit": "img/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus, args.conv_math, args.conv or "tc3", args.rng),      # the SAME config as the afan arm's line
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(_r(line, 5), separators=(",", ":")), flush=True)


# ---------------------------------------------------------------------------------------------------
# kernel rooflines: the hand-written kernels at the workload's shapes, L2-cold
# ---------------------------------------------------------------------------------------------------
L2_BYTES = 126 * 1024 * 1024


def kernel_rooflines(pkg, dev, reps=10, in_step_only=False):
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
    # tensor-pipe denominator: the driver measures dense bf16 (cuBLAS); kind::tf32 runs at half that rate
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tensor_peak = float(json.load(f)["bf16_tflops"]) / 2
    except Exception:
        tensor_peak = 1590.0 / 2          # fallback (B200_PROFILING.md)

    def measure(name, bytes_per_launch, footprint, make_set, run, launches_per_iter, note, flops=None, bound="fp32_ffma"):
        if in_step_only and launches_per_iter == 0:          # N > 1: only the shapes the timed step launches
            return
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
            pk = ffma_peak if bound == "fp32_ffma" else tensor_peak
            row.update({"bound": bound, "flops": flops, "achieved": flops / t / 1e12, "peak": pk, "unit": "TFLOP/s",
                        "frac": flops / t / 1e12 / pk, "hbm_gbs": bytes_per_launch / t / 1e9})
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
                                      ("G2 256x64x56x56", (2, 256, 64, 56, 56), 0),
                                      # BASELINE config 5 (DeepLabv3+ R101, 4 x 513^2): ASPP input / ASPP branch / decoder
                                      ("cfg5 G1 4x2048x33x33", (1, 4, 2048, 33, 33), 0),
                                      ("cfg5 G2 4x256x33x33", (2, 4, 256, 33, 33), 0),
                                      ("cfg5 G2 4x256x129x129", (2, 4, 256, 129, 129), 0)):
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
        umma = pkg.conv.MODE == "tc3" and ops.conv3x3_umma_supported(N, C, H)
        math = "umma" if umma else {"afan": "fp32", "tf32": "tf32", "3xtf32": "3xtf32"}.get(pkg.conv.MODE, "fp32")
        if umma:
            measure(f"conv3x3 tcgen05 fwd/dgrad [{N}x{C}x{H}x{H}]", 8 * E + 4 * 9 * C * C, 8 * E, mk,
                    lambda t: ops.conv3x3(t["x"], t["wf"], math="umma"), lpi_conv,
                    "3x3 s1 p1 conv as a tcgen05 implicit GEMM (kind::tf32, 3xTF32 split, TMEM accumulators): 18*C FLOP per "
                    "output element (the split issues 3x that on the tensor pipe); read x + write y = 8 B/elem",
                    flops=fl_conv, bound="tensor")
        else:
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
# parity evidence carried by the bench line itself
# ---------------------------------------------------------------------------------------------------
def multi_gpu_parity(pkg, dev, pg, rank, world, exchange, graph, sync_bn):
    """world > 1, before timing: 2 iterations of a sharded step (this run's exchange / graph mode) must (a) leave
    bit-identical weights and BatchNorm buffers on every rank and (b) reproduce the SINGLE-process step at the GLOBAL
    batch (the multi-GPU oracle; replaces nn.DataParallel of Segmentation/main_aug_final.py:119,131)."""
    dist = torch.distributed
    per_rank, iters = 32, 2
    kw = dict(perturb_idx=5, steps=2, gamma=1.0, eps=2.0, randinit=True, clip=True)
    g = torch.Generator().manual_seed(21)
    B = per_rank * world
    images = [torch.rand(B, 3, 32, 32, generator=g) for _ in range(iters)]
    targets = [torch.randint(0, 10, (B,), generator=g) for _ in range(iters)]
    noises = [torch.rand(B, 16, 32, 32, generator=g) for _ in range(iters)]
    torch.manual_seed(3)
    model = pkg.resnet_s.ResNet(num_blocks=(1, 2, 2)).to(dev)       # second block of stages 2 / 3: the folded-BatchNorm path
    init = {k: v.clone() for k, v in model.state_dict().items()}
    folded0 = pkg.resnet_s.fused_forward_calls
    tr = pkg.trainer.AfanTrainer(model, process_group=pg, sync_bn=sync_bn, use_cuda_graph=graph, bn_exchange=exchange, **kw)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    losses = []
    for i in range(iters):
        out = tr.step(images[i][sl].to(dev), targets[i][sl].to(dev), noises[i][sl].to(dev))
        l = out["loss"].detach().clone()
        dist.all_reduce(l)
        losses.append(float(l) / world)
    tr.check()
    sharded = {k: v.detach().clone() for k, v in model.state_dict().items()}
    flat = torch.cat([v.reshape(-1).double() for v in sharded.values()])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool((lo == hi).all())
    tr.close()
    res = {"ranks_bit_identical": identical, "exchange": tr.bn_exchange_used, "graph": graph, "per_rank_batch": per_rank,
           "folded_block_calls": pkg.resnet_s.fused_forward_calls - folded0}
    if sync_bn:                       # per-replica statistics (--no-sync-bn) have no single-process equivalent
        ref_model = pkg.resnet_s.ResNet(num_blocks=(1, 2, 2)).to(dev)
        ref_model.load_state_dict(init)
        ref = pkg.trainer.AfanTrainer(ref_model, use_cuda_graph=False, **kw)
        ref_losses = [float(ref.step(images[i].to(dev), targets[i].to(dev), noises[i].to(dev))["loss"]) for i in range(iters)]
        single = ref_model.state_dict()
        ref.close()
        worst, ok = ("", 0.0), True
        for k, v in single.items():
            if not v.dtype.is_floating_point:
                ok &= bool((sharded[k] == v).all())
                continue
            err = float((sharded[k] - v).abs().max())
            if err > worst[1]:
                worst = (k, err)
            ok &= err <= 2e-4 + 2e-3 * float(v.abs().max())
        loss_rel = max(abs(a - b) / abs(b) for a, b in zip(losses, ref_losses))
        ok &= loss_rel <= 1e-4
        res.update({"vs_global_batch_ok": ok, "loss_rel_err": loss_rel, "worst_abs_err": worst[1], "worst_key": worst[0],
                    "losses": losses, "ref_losses": ref_losses})
    flag = torch.tensor([1 if (identical and res.get("vs_global_batch_ok", True)) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["status"] = "ok" if int(flag) else "FAIL"
    return res


def parity_iter0(pkg, dev):
    """N = 1: iteration 0 of the benchmarked workload (full size, injected random start) next to the CPU port of the
    reference iteration on the same weights / inputs / noise."""
    from oracle import afan_ref_torch as ref_t
    w = WORKLOAD
    torch.manual_seed(3)
    model = pkg.resnet_s.ResNet(num_blocks=w["num_blocks"], num_classes=w["num_classes"])
    ref = ref_t.CifarResNetRef(w["num_blocks"], w["num_classes"])
    ref.load_state_dict(model.state_dict())
    model.to(dev)
    n = w["batch_per_gpu"]
    g = torch.Generator().manual_seed(11)
    x, y = torch.rand(n, *w["image"], generator=g), torch.randint(0, w["num_classes"], (n,), generator=g)
    noise = torch.rand(n, 16, 32, 32, generator=g)
    kw = dict(steps=w["steps"], gamma=w["gamma"], eps=w["eps"], perturb_idx=w["perturb_idx"], randinit=True, clip=True)
    tr = pkg.trainer.AfanTrainer(model, use_cuda_graph=False, **kw)
    out = tr.step(x.to(dev), y.to(dev), noise.to(dev))
    got = float(out["loss"])
    linf = float(out["linf"].max())
    tr.close()
    ref.train()
    opt, crit = ref_t.make_sgd(ref), torch.nn.CrossEntropyLoss()
    torch.set_num_threads(os.cpu_count() or 1)
    loss_ref, _, _, linf_ref, _ = ref_t.afan_train_iteration(ref, opt, crit, x, y, noise=noise, **kw)
    return {"loss": round(got, 6), "port_loss": round(float(loss_ref), 6), "linf_max": linf, "port_linf_max": float(linf_ref.max())}


def _r(v, nd=4):
    """Round floats for the compact line."""
    if isinstance(v, float):
        return float(f"{v:.{nd}g}") if abs(v) < 1e4 else round(v, 1)
    if isinstance(v, dict):
        return {k: _r(x, nd) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_r(x, nd) for x in v]
    return v


def emit(line, detail, args, world):
    """Full record -> file (gpurun_out/ is scratch; copy to profiles/ to keep); ONE compact JSON line -> stdout (last)."""
    path = args.detail_file or os.path.join(ROOT, "gpurun_out", f"bench_detail_n{world}.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump({**line, **detail}, f, indent=1)
        line["detail_file"] = os.path.relpath(path, ROOT)
    except OSError as e:
        line["detail_file"] = f"unwritable: {e}"[:60]
    text = json.dumps(_r(line, 4), separators=(",", ":"))
    if len(text) > 1400:                  # the driver parses the tail of stdout: never let the line outgrow it
        for k in ("afan_kernels_per_step", "variants_ms_per_step", "multi_gpu_parity_err", "parity_iter0", "reference_on_gpu_ms",
                  "roofline_hbm"):
            if k in line and len(text) > 1400:
                line.pop(k)
                text = json.dumps(_r(line, 4), separators=(",", ":"))
    sys.stdout.flush()
    print(text, flush=True)


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
    ap.add_argument("--conv", default=None, choices=["afan", "cudnn", "3xtf32", "tc3"],
                    help="3x3 tail convolutions: hand-written sm_100a kernels (afan = strict-fp32 FFMA, or the mma.sync TF32 "
                         "twin under --conv-math tf32; tc3 = tcgen05 3xTF32 implicit GEMM), the cuDNN library path, or the "
                         "mma.sync 3xTF32 split kernels.  Default: the package default (AFAN_CONV)")
    ap.add_argument("--rng", default="philox", choices=["philox", "injected"],
                    help="random start: on-device Philox (fast path) or noise injected from the host generator (the reference's)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-variants", action="store_true", help="do not time the TF32 / cuDNN convolution variants of the step")
    ap.add_argument("--skip-rooflines", action="store_true")
    ap.add_argument("--skip-parity", action="store_true", help="skip the in-bench parity legs (multi_gpu_parity / parity_iter0)")
    ap.add_argument("--detail-file", default=None, help="where the full record (kernel table, families, ...) is written")
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
    if args.conv is None:
        args.conv = pkg.conv.MODE if pkg.conv.MODE in ("afan", "cudnn", "3xtf32", "tc3") else "tc3"
    pkg.conv.MODE = "tf32" if (args.conv == "afan" and args.conv_math == "tf32") else args.conv
    w = WORKLOAD
    detail = {}

    # (0) parity legs, before any timing
    parity = None
    if world > 1 and not args.skip_parity:
        parity = multi_gpu_parity(pkg, dev, pg, rank, world, args.bn_exchange, not args.no_graph, not args.no_sync_bn)
        detail["multi_gpu_parity_detail"] = parity
        if rank == 0:
            print("MULTI_GPU_CHECK " + json.dumps(_r(parity, 4)), file=sys.stderr, flush=True)

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
    host_u = [torch.rand(n, 16, 32, 32, generator=g).pin_memory() for _ in range(4)] if args.rng == "injected" else None
    dev_x, dev_y = [t.to(dev) for t in host_x], [t.to(dev) for t in host_y]
    dev_u = [t.to(dev) for t in host_u] if host_u else None

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
    trainer.step(dev_x[0], dev_y[0], dev_u[0] if dev_u else None)       # builds arena, captures the graph
    out = {}

    def step_dev(i):
        out["r"] = trainer.step(dev_x[i % 4], dev_y[i % 4], dev_u[i % 4] if dev_u else None)
    sec, clocks = timed_run(step_dev)
    loss_dev = float(out["r"]["loss"])

    # (2) end to end through the public pipeline (prefetch.DevicePrefetcher, as main_perturb.py uses it): pinned host batch i+1
    #     -> H2D on the copy stream while step i runs -> step -> D2H loss + host sync EVERY step.  Every batch's copy is issued
    #     and completed inside the timed region (the first one before the first step; one batch is in flight at its end).
    host_loss = [torch.zeros(1).pin_memory() for _ in range(2)]
    read_done = [torch.cuda.Event() for _ in range(2)]
    seen = {"last": None, "losses": 0}

    def host_batches():
        i = 0
        while True:
            yield host_x[i % 4], host_y[i % 4], (host_u[i % 4] if host_u else None)
            i += 1
    feed = {"it": None}

    def step_e2e(i):
        if feed["it"] is None:
            feed["it"] = iter(pkg.prefetch.DevicePrefetcher(host_batches(), dev))
        dx, dy, du = next(feed["it"])
        r = trainer.step(dx, dy, du)
        k = seen["losses"] % 2
        host_loss[k].copy_(r["loss"].reshape(1), non_blocking=True)      # D2H of THIS step's loss, every step
        read_done[k].record()
        if seen["last"] is not None:                                      # the host reads each loss one step late, while the next
            read_done[seen["last"]].synchronize()                         # step is already running (asynchronous logging; the
            float(host_loss[seen["last"]])                                # reference prints every print_freq steps, main_perturb.py:208)
        seen["last"] = k
        seen["losses"] += 1
    sec_e2e, _ = timed_run(step_e2e)
    torch.cuda.synchronize()
    loss_e2e_last = float(host_loss[seen["last"]])           # the last step's loss: its copy was issued inside the timed region

    def step_e2e_strict(i):                                  # same pipeline, but the host blocks on every step's loss before the next
        dx, dy, du = next(feed["it"])
        r = trainer.step(dx, dy, du)
        host_loss[0].copy_(r["loss"].reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    sec_e2e_strict, _ = timed_run(step_e2e_strict)
    trainer.check()                                          # a timed-out statistics exchange must fail the run, not pass silently

    per_iter = trainer.kernel_launches_per_iter      # afan kernels per iteration, counted at capture/trace time

    global_batch = n * world
    h2d = (host_x[0].numel() * 4 + host_y[0].numel() * 8 + (host_u[0].numel() * 4 if host_u else 0)) * world
    line = {"metric": "A-FAN train img/s", "value": global_batch * args.steps / sec, "unit": "img/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(world, args.conv_math, args.conv, args.rng), "clocks": clocks,
            "e2e": {"value": global_batch * args.steps / sec_e2e, "unit": "img/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": 1e3 * sec_e2e / args.steps,
                    "blocking_read_ms_per_step": 1e3 * sec_e2e_strict / args.steps,
                    "pipeline": "H2D of batch i+1 on a copy stream; loss D2H every step, read by the host one step late"},
            "gpu_launches": per_iter * args.steps, "afan_kernels_per_step": per_iter}
    detail.update({"bn1_folded_into_conv2_in_ascent": bool(pkg.resnet_s.FUSE_BN1) and world == 1 and pkg.conv.MODE == "tc3",
                   "wgrad_tcgen05_c32": bool(pkg.conv.WGRAD_UMMA) and pkg.conv.MODE == "tc3",
                   "cuda_graph": not args.no_graph, "sync_bn": (not args.no_sync_bn) and world > 1,
                   "bn_exchange": trainer.bn_exchange_used, "final_loss": loss_dev})
    if world > 1:
        line["bn_exchange"] = trainer.bn_exchange_used
        if parity is not None:
            line["multi_gpu_parity"] = parity["status"]
            if "worst_abs_err" in parity:
                line["multi_gpu_parity_err"] = {"loss_rel": parity["loss_rel_err"], "weights_abs": parity["worst_abs_err"],
                                                "ranks_bit_identical": parity["ranks_bit_identical"]}

    # other convolution paths of the SAME step, for context only (never the headline): single GPU, device-resident inputs
    if world == 1 and not args.skip_variants and args.conv_math == "fp32" and not args.no_graph:
        variants = {}
        det = torch.backends.cudnn.deterministic
        head_mode = pkg.conv.MODE
        cands = (("afan_fp32_ffma", "afan", False), ("afan_tcgen05_3xtf32", "tc3", False), ("afan_mma_sync_tf32", "tf32", True),
                 ("cudnn_fp32_nondet", "cudnn", False), ("cudnn_tf32_nondet", "cudnn", True))
        for vname, mode, tf32 in cands:
            if mode == head_mode and not tf32:
                continue
            if mode == "tc3" and not getattr(pkg.ops, "CONV_TC_AVAILABLE", False):
                continue
            pkg.conv.MODE = mode
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            # the library path is timed with its fastest (non-deterministic) algorithms
            torch.backends.cudnn.deterministic = det and mode != "cudnn"
            torch.manual_seed(3)
            vm = pkg.resnet_s.ResNet(num_blocks=w["num_blocks"], num_classes=w["num_classes"]).to(dev)
            vt = pkg.trainer.AfanTrainer(vm, perturb_idx=w["perturb_idx"], steps=w["steps"], gamma=w["gamma"], eps=w["eps"],
                                         randinit=w["randinit"], clip=w["clip"], rng="philox", seed=3, use_cuda_graph=True)
            vt.step(dev_x[0], dev_y[0])
            vsec, _ = timed_run(lambda i: vt.step(dev_x[i % 4], dev_y[i % 4]))
            variants[vname] = 1e3 * vsec / args.steps
            vt.close()
            del vm, vt
        pkg.conv.MODE = head_mode
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.deterministic = det
        line["variants_ms_per_step"] = variants

    if world > 1 and not args.skip_variants and not args.no_sync_bn and not args.no_graph:
        # context for the scaling number: the SAME sharded step with per-replica BatchNorm statistics -- what the reference's
        # own multi-GPU mode computes (nn.DataParallel, Segmentation/main_aug_final.py:119,131) -- i.e. the step without the
        # ~390 per-layer statistics exchanges the global-batch semantics costs.  Not the headline.
        torch.manual_seed(3)
        vm = pkg.resnet_s.ResNet(num_blocks=w["num_blocks"], num_classes=w["num_classes"]).to(dev)
        vt = pkg.trainer.AfanTrainer(vm, perturb_idx=w["perturb_idx"], steps=w["steps"], gamma=w["gamma"], eps=w["eps"],
                                     randinit=w["randinit"], clip=w["clip"], rng="philox", seed=3 + rank, process_group=pg,
                                     sync_bn=False, use_cuda_graph=True, bn_exchange=args.bn_exchange)
        vt.step(dev_x[0], dev_y[0])
        vsec, _ = timed_run(lambda i: vt.step(dev_x[i % 4], dev_y[i % 4]))
        detail["per_replica_bn_ms_per_step"] = 1e3 * vsec / args.steps
        detail["per_replica_bn_img_per_s"] = global_batch * args.steps / vsec
        vt.close()
        del vm, vt

    if rank == 0 and not args.skip_rooflines:
        ks, peak_src, ffma_peak = kernel_rooflines(pkg, dev, in_step_only=world > 1)
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tmap = json.load(f)
        except Exception:
            tmap = {}
        for k in ks:
            k["traffic"] = tmap.get(k["kernel"])
        # Roofline per kernel FAMILY (all in-step shapes of a kernel, weighted by their launches in one step):
        # achieved = algorithmic bytes (HBM-bound families) or FLOPs (the convolutions) per step / device time
        # per step.  `roofline` = the family with the largest share of the step; `roofline_hbm` = the largest HBM-bound one.
        step_us = 1e3 * sec / args.steps * 1e3
        fam = {}
        for k in ks:
            if k["launches_per_iter"] > 0:
                f = fam.setdefault(k["family"], {"bound": k["bound"], "peak": k["peak"], "unit": k["unit"], "work": 0.0, "bytes": 0.0,
                                                 "us": 0.0, "launches": 0, "traffic": 0.0, "traffic_ok": True, "shapes": []})
                f["work"] += (k["flops"] if k["bound"] != "hbm" else k["bytes"]) * k["launches_per_iter"]
                f["bytes"] += k["bytes"] * k["launches_per_iter"]
                f["us"] += k["us"] * k["launches_per_iter"]
                f["launches"] += k["launches_per_iter"]
                f["shapes"].append(k["kernel"])
                if k["traffic"] is None:
                    f["traffic_ok"] = False
                else:
                    f["traffic"] += k["traffic"] * k["launches_per_iter"]

        def fam_line(name, f, full=False):
            scale = 1e9 if f["bound"] == "hbm" else 1e12
            achieved = f["work"] / (f["us"] * 1e-6) / scale
            d = {"kernel": name, "bound": f["bound"], "achieved": achieved, "peak": f["peak"], "unit": f["unit"],
                 "frac": achieved / f["peak"], "traffic": f["traffic"] / f["launches"] if f["traffic_ok"] else None,
                 "algorithmic_bytes": f["bytes"] / f["launches"], "launches_per_step": f["launches"],
                 "avg_us": f["us"] / f["launches"], "share_of_step": f["us"] / step_us}
            if not full:
                d.pop("launches_per_step"); d.pop("avg_us")
            if full:
                d["shapes"] = f["shapes"]
                d["peak_source"] = peak_src if f["bound"] == "hbm" else k_peak_note(f["bound"])
            return d
        dom_name, dom = max(fam.items(), key=lambda kv: kv[1]["us"])
        line["roofline"] = fam_line(dom_name, dom)
        hbm = {nm: f for nm, f in fam.items() if f["bound"] == "hbm"}
        if hbm and dom["bound"] != "hbm":
            hn, hf = max(hbm.items(), key=lambda kv: kv[1]["us"])
            line["roofline_hbm"] = fam_line(hn, hf)
        detail["roofline_timing"] = ("per shape: CUDA events around a graph of back-to-back launches over rotating tensor sets (4x L2, "
                                     "every launch L2-cold), weighted by the launches of that shape in one step; traffic / "
                                     "algorithmic_bytes are per-launch averages")
        detail["roofline_families"] = {nm: fam_line(nm, f, full=True) for nm, f in fam.items()}
        detail["ffma_peak_tflops_measured"] = ffma_peak
        detail["kernels"] = ks
    if world > 1:
        torch.distributed.barrier()
        if not args.no_sync_bn:
            detail["bn_fwd_us_per_call_G1_128x32x16x16"] = exchange_microbench(pkg, dev, trainer.mailbox, pg)
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        try:
            r32, rtf = gpu_reference(dev, tf32=False), gpu_reference(dev, tf32=True)
            detail["reference_on_gpu"] = {"strict_fp32": r32, "stock_tf32_conv": rtf}
            line["reference_on_gpu_ms"] = {"fp32": r32["ms_per_step"], "tf32_default": rtf["ms_per_step"]}
        except Exception as exc:                     # context number only: never fail the bench line over it
            detail["reference_on_gpu"] = {"error": repr(exc)[:200]}
        if not args.skip_parity:
            try:
                line["parity_iter0"] = parity_iter0(pkg, dev)
            except Exception as exc:
                line["parity_iter0"] = {"error": repr(exc)[:120]}
        cb = cpu_reference(steps=4, warmup=1)
        line["cpu_baseline"] = {"value": cb["value"], "unit": cb["unit"], "cores": cb["cores"], "kind": cb["kind"],
                                "sample": f"4 iterations of this workload, batch {n}"}
        detail["cpu_baseline_detail"] = cb
    if rank == 0:
        emit(line, detail, args, world)
    if world > 1:
        trainer.close()                      # a live graph with captured NCCL ops blocks communicator teardown
        torch.distributed.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0 if (parity is None or parity["status"] == "ok") else 3)


def k_peak_note(bound):
    if bound == "fp32_ffma":
        return ("fp32 FFMA rate measured live by afan_ffma_probe (8x8 outer-product loop, 2 CTAs x 256 threads per SM) on this "
                "device at its current clocks; nominal 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s")
    return ("tensor: MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 runs at half the bf16 rate; no measured TF32 peak exists); "
            "achieved counts the convolution's algorithmic FLOPs once, the 3xTF32 split issues 3x that")


if __name__ == "__main__":
    main()
