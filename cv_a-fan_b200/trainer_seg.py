"""A-FAN training step of the Segmentation flavour (row f3): Segmentation/main_aug_final.py:160-232 on B200.

Reference iteration (per batch, DeepLabv3+, train mode):
    #1 head forward to backbone stage `se`           -> feature_map_se (detached), low_level (graph kept)   :165,173
    #2 FULL forward + ASPP / concat ('<sd>_head')     -> decoder feature dict (graph kept)                   :166,170-171
    PGD on the stage-`se` feature                     (steps x tail forward + dgrad)                         :176-188
    decoder_PGD on the ASPP / concat feature          (steps x decoder forward + dgrad)                      :190-201
    mix_feature / uniform noise on the decoder point, 3 SAT points on the clean->adv segment (+ per-point mix) :203-211
    #3 FULL clean forward, two stage-`se` tails, one decoder tail                                             :213-221
    loss = 0.7 l0 + 0.1 l1 + 0.1 l2 + 0.1 l3, backward, SGD (backbone lr x 0.1)                               :223-231

B200 design.
  * head cache: forwards #1, #2 and #3 see the same images in train mode, i.e. they compute the same activations three
    times.  Here the backbone and the ASPP run ONCE; the graph of that sweep serves all three uses (autograd sums the
    gradients exactly as three graphs would), and the BatchNorm running statistics of the shared layers get the
    closed-form k-fold update the reference's repeated passes produce (flat statistic arenas: 2 launches).
    `head_cache=False` replays the reference schedule literally.
  * every elementwise span of the ascents is the fused PGD kernel (`segmentation.PGD` / `decoder_PGD`), the SAT
    points + their `mix_feature` are ONE launch (`segmentation.sat_sample_points`), random starts come from the
    in-register Philox generator unless `noise=` injects the reference's CPU draws.
  * parameters live in two flat arenas (backbone / classifier: the reference's two SGD groups) with one fused SGD
    launch each; learning rates sit in device memory (`set_lr`, PolyLR-ready).
  * dual BN: the tail's BatchNorm layers (stages after `se`, ASPP, decoder) run on the hand-written dual-BN kernels and
    the two stage-`se` tail passes of the final forward are ONE batched pass with two statistic groups (`dual_bn=True`).
Convolutions of DeepLab and the head's BatchNorm stay library kernels (out of the hand-written scope).

Declared deviations from the reference schedule: with the head cache (a) the shared layers' running statistics
get the k-fold update up front, where the reference interleaves clean pass #2, the ascents' tail updates and clean pass
#3 -- an exponential average depends on that order, the statistics agree to 2e-2 (tested); (b) ASPP's Dropout draws ONE
mask for the decoder anchor and `out0` where the reference's two clean passes draw two; (c) the flat SGD arena applies
weight decay to every parameter of the group, torch.optim.SGD skips parameters whose .grad is None (none exists in this
model: every parameter receives a gradient each iteration).  `head_cache=False` replays the reference schedule literally.
"""
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import dual_bn as dual_bn_mod
from . import ops, segmentation
from ._lib import AfanError
from .dual_bn import DualBatchNorm2d

_BN = (nn.BatchNorm2d, DualBatchNorm2d)         # (GroupedLibraryBatchNorm2d is an nn.BatchNorm2d)


class _Arena:
    """Flat fp32 parameter / gradient / momentum buffers for one SGD group."""

    def __init__(self, params, lr, device):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        pad = (-n) % 4
        self.param = torch.zeros(n + pad, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.param)
        self.buf = torch.zeros_like(self.param)
        off = 0
        for p in self.params:
            k = p.numel()
            self.param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.param[off:off + k].view_as(p.data)
            p.grad = self.grad[off:off + k].view_as(p.data)
            off += k
        self.lr = torch.full((1,), float(lr), dtype=torch.float32, device=device)


class _StatArena:
    """running_mean / running_var of a set of BatchNorm layers as views of ONE flat tensor, so that the k-fold update of
    a batch the reference forwards k times is two launches: new = a * after_one + b * before."""

    def __init__(self, bns, device):
        self.bns = list(bns)
        n = sum(2 * m.num_features for m in self.bns)
        self.flat = torch.zeros(max(n, 1), dtype=torch.float32, device=device)
        off = 0
        for m in self.bns:
            c = m.num_features
            for name in ("running_mean", "running_var"):
                view = self.flat[off:off + c]
                view.copy_(getattr(m, name))
                setattr(m, name, view)                     # buffers stay registered under the same names
                off += c
        self.momentum = self.bns[0].momentum if self.bns else 0.1
        if any(m.momentum != self.momentum for m in self.bns):
            raise AfanError("a statistic arena needs one momentum")
        self.before = torch.empty_like(self.flat)

    def snapshot(self):
        self.before.copy_(self.flat)

    def replay(self, k: int):
        """The layers were just updated ONCE with this batch; make it k times (same batch statistics every time)."""
        if k <= 1 or not self.bns:
            return
        m = self.momentum
        a = (1.0 - (1.0 - m) ** k) / m
        b = (1.0 - m) ** k - a * (1.0 - m)
        # through .data: like the in-kernel update of F.batch_norm itself, this must not bump the version counter of the
        # buffers autograd saved for the pending backward
        self.flat.data.mul_(a).add_(self.before, alpha=b)
        for bn in self.bns:
            bn.num_batches_tracked.data += k - 1


class SegAfanTrainer:
    def __init__(self, model: nn.Module, *, pertub_idx_se: int = 3, pertub_idx_sd: str = "aspp", steps: int = 1,
                 eps: float = 2.0, gamma_se: float = 0.5, gamma_sd: float = 0.5, randinit: bool = False,
                 clip: bool = False, mix_sd: bool = False, noise_sd: float = 0.0, mix_layer: str = "00",
                 lr: float = 0.01, momentum: float = 0.9, weight_decay: float = 1e-4,
                 criterion: Optional[nn.Module] = None, head_cache: bool = True, rng: str = "philox", seed: int = 0,
                 dual_bn: bool = True, use_cuda_graph: bool = False):
        """Flag names / units follow Segmentation/args.py:19-40 (eps, gamma_* in 1/255).

        dual_bn: the BatchNorm layers of the TAIL (backbone stages after `pertub_idx_se`, ASPP, decoder:
        Segmentation/network/backbone/resnet.py:127,144, _deeplab.py:47-80) become DualBatchNorm2d (the hand-written
        sm_100a kernels), and the two stage-`se` tail passes of the final forward (main_aug_final.py:222-223) run as ONE
        batched pass with two statistic groups -- the same per-pass statistics, shared affine, running average updated in
        pass order."""
        if pertub_idx_se not in (1, 2, 3, 4) or pertub_idx_sd not in ("aspp", "concat"):
            raise AfanError("pertub_idx_se must be 1..4 and pertub_idx_sd 'aspp' or 'concat'")
        if len(mix_layer) != 2 or any(ch not in "01" for ch in mix_layer):
            raise AfanError("mix_layer is two 0/1 flags, e.g. '01' (main_aug_final.py:26-27)")
        self.model, self.se, self.sd = model, pertub_idx_se, pertub_idx_sd
        self.steps, self.eps = int(steps), eps / 255.0
        self.gamma_se, self.gamma_sd = gamma_se / 255.0, gamma_sd / 255.0
        self.gamma_sd_raw = float(gamma_sd)                 # :206-207 scale the uniform noise by opts.gamma_sd itself
        self.randinit, self.clip, self.mix_sd, self.noise_sd = randinit, clip, mix_sd, float(noise_sd)
        self.f0, self.f1 = int(mix_layer[0]), int(mix_layer[1])
        self.momentum, self.weight_decay = momentum, weight_decay
        self.criterion = criterion if criterion is not None else nn.CrossEntropyLoss(ignore_index=255, reduction="mean")
        self.head_cache, self.rng, self.seed = head_cache, rng, int(seed)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise AfanError("SegAfanTrainer needs the model on a CUDA device: there is no CPU path")
        self.dual_bn = bool(dual_bn)
        if self.dual_bn:
            from torchvision.models.segmentation.deeplabv3 import ASPPPooling
            for k in range(pertub_idx_se + 1, 5):
                dual_bn_mod.convert_batchnorm(getattr(model.backbone, f"layer{k}"))
            dual_bn_mod.convert_batchnorm(model.classifier, pooled=(ASPPPooling,))     # image-pooling branch: library kernel
        # the reference's two SGD groups (main_aug_final.py:79-82)
        self.arenas = [_Arena(model.backbone.parameters(), 0.1 * lr, self.device),
                       _Arena(model.classifier.parameters(), lr, self.device)]
        self._lr = float(lr)
        bb = model.backbone
        shared = [bb.bn1] + [m for k in range(1, self.se + 1) for m in getattr(bb, f"layer{k}").modules()
                             if isinstance(m, _BN)]
        rest = [m for k in range(self.se + 1, 5) for m in getattr(bb, f"layer{k}").modules() if isinstance(m, _BN)]
        aspp = [m for m in model.classifier.aspp.modules() if isinstance(m, _BN)]
        if self.sd == "concat":           # '<concat>_head' (#2) and the clean pass (#3) both run the low-level projection
            aspp += [m for m in model.classifier.project.modules() if isinstance(m, _BN)]
        # (statistic arena, how often the reference forwards these layers on the CLEAN images per iteration)
        self._replays = [(_StatArena(shared, self.device), 3), (_StatArena(rest, self.device), 2),
                         (_StatArena(aspp, self.device), 2)] if head_cache else []
        self.rng_offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.iterations = 0
        # use_cuda_graph: the whole iteration (head sweep, both ascents, four training forwards, backward, both SGD launches)
        # is captured once and replayed -- the eager iteration is HOST-bound at config 5 (45.8 ms to enqueue ~4500 launches
        # against 47.6 ms of device time, profiles/probes/host_vs_device.py).  Needs rng='philox' (device-side offsets) or
        # injected noise; the learning rate already lives on the device.
        self.use_graph = bool(use_cuda_graph)
        if self.use_graph and self.randinit and self.rng != "philox":
            raise AfanError("use_cuda_graph needs rng='philox' (the CPU random stream of rng='reference' cannot be replayed)")
        self._graph, self._static = None, {}
        self._dual = [m for m in model.modules() if isinstance(m, DualBatchNorm2d)]

    def set_lr(self, lr: float):
        """Base learning rate (PolyLR, utils/scheduler.py:3-12, is applied by the caller): backbone gets 0.1 x."""
        if lr != self._lr:
            self.arenas[0].lr.fill_(0.1 * lr)
            self.arenas[1].lr.fill_(lr)
            self._lr = float(lr)

    # ---------------------------------------------------------------------------------------------
    def _extras(self, noise, key, numel):
        if not self.randinit:
            return {}
        if noise is not None and key in noise:
            return {"noise": noise[key]}
        if self.rng == "philox":
            ex = {"rng": "philox", "seed": self.seed, "offset_device": self.rng_offset.clone()}
            self.rng_offset += (numel + 3) // 4
            return ex
        return {}                                              # rng='reference': CPU torch.rand like the reference

    def _iteration(self, images, labels, noise):
        model, crit, se, sd = self.model, self.criterion, self.se, self.sd
        size = images.shape[-2:]
        if self.head_cache:
            for arena, _ in self._replays:
                arena.snapshot()
            st = model.backbone.stages(images, 4)                                   # ONE sweep for #1, #2 and #3
            low, out4 = st[1], st[4]
            aspp_out = model.classifier.aspp(out4)
            dec = model.classifier.concat(low, aspp_out) if sd == "concat" else aspp_out
            for arena, k in self._replays:
                arena.replay(k)
            feat_se = st[se].detach()
            dec_dict = {"low_level": low, "out": out4, "adv": dec}
            low_for_ascent = low.detach()
        else:
            out_se = model({"x": images, "adv": None, "out_idx": se, "flag": "head"})                        # :165
            dec_dict = model({"x": images, "adv": None, "out_idx": sd + "_head", "flag": "clean"})           # :166
            low, feat_se = out_se["low_level"], out_se["out"].detach()
            low_for_ascent = low
        feat_sd = dec_dict["adv"].detach()

        adv_se = segmentation.PGD(feat_se, images, low_for_ascent, crit, y=labels, model=model, steps=self.steps,
                                  eps=self.eps, gamma=self.gamma_se, idx=se, randinit=self.randinit, clip=self.clip,
                                  **self._extras(noise, "se", feat_se.numel()))                              # :176-188
        asc_dict = dict(dec_dict)
        if self.head_cache:
            asc_dict["low_level"] = low_for_ascent
        asc_dict = segmentation.decoder_PGD(asc_dict, images, crit, y=labels, model=model, steps=self.steps, eps=self.eps,
                                            gamma=self.gamma_sd, idx=sd, randinit=self.randinit, clip=self.clip,
                                            **self._extras(noise, "sd", feat_sd.numel()))                    # :190-201
        adv_sd = asc_dict["adv"].detach()
        if self.mix_sd:
            adv_sd = segmentation.mix_feature(feat_sd, adv_sd)                                               # :204-205
        if self.noise_sd != 0:
            u = noise["noise_sd"].to(self.device) if (noise is not None and "noise_sd" in noise) else \
                torch.rand(adv_sd.shape, device=self.device)
            adv_sd = adv_sd + (2.0 * u - 1.0) * self.gamma_sd_raw * self.noise_sd                           # :206-207 (opts.gamma_sd, not /255)
        dec_dict["adv"] = adv_sd
        pts = segmentation.sat_sample_points(feat_se, adv_se, 3, mix=[self.f0, self.f1])                     # :210-214

        if self.head_cache:
            feats = {"low_level": low, "out": out4}
            if sd == "concat":
                out0 = model.logits({"adv": dec}, size, return_type="concat_tail")
            else:
                out0 = model.logits({"low_level": low, "adv": aspp_out}, size, return_type="aspp_tail")
            del feats
        else:
            out0 = model({"x": images, "adv": None, "out_idx": 0, "flag": "clean"})                          # :221
        if self.dual_bn:
            # :222-223 as ONE pass over [point 1; point 2] with two BatchNorm statistic groups ("dual BN")
            nb = images.shape[0]
            with dual_bn_mod.statistic_groups(2):
                out12 = model({"x": images, "adv": torch.cat([pts[1], pts[2]], 0), "out_idx": se, "flag": "tail",
                               "low_level_feat": torch.cat([low, low], 0)})
            out1, out2 = out12[:nb], out12[nb:]
        else:
            out1 = model({"x": images, "adv": pts[1], "out_idx": se, "flag": "tail", "low_level_feat": low})     # :222
            out2 = model({"x": images, "adv": pts[2], "out_idx": se, "flag": "tail", "low_level_feat": low})     # :223
        out3 = model({"x": images, "adv": dec_dict, "out_idx": sd + "_tail", "flag": "clean"})               # :224
        l0, l1, l2, l3 = crit(out0, labels), crit(out1, labels), crit(out2, labels), crit(out3, labels)
        loss = 0.7 * l0 + 0.1 * l1 + 0.1 * l2 + 0.1 * l3                                                     # :233
        return loss, torch.stack([l0.detach(), l1.detach(), l2.detach(), l3.detach()])

    def _eager(self, images, labels, noise):
        for a in self.arenas:
            a.grad.zero_()                                                                                   # :162
        loss, parts = self._iteration(images, labels, noise)
        loss.backward()                                                                                      # :235
        for a in self.arenas:                                                                                # :236
            ops.sgd_momentum_(a.param, a.grad, a.buf, a.lr, momentum=self.momentum, weight_decay=self.weight_decay)
        return loss.detach(), parts

    def _capture(self, images, labels, noise):
        st = self._static
        st["images"], st["labels"] = images.clone(), labels.clone()
        st["noise"] = {k: v.clone() for k, v in noise.items()} if noise else None
        for m in self._dual:
            m.flush_batches_tracked()
        tensors = dict(self.model.named_parameters())
        tensors.update({"buffer:" + k: v for k, v in self.model.named_buffers()})
        snap = {"model": {k: v.detach().clone() for k, v in tensors.items()},
                "arena": [(a.param.clone(), a.buf.clone()) for a in self.arenas], "rng": self.rng_offset.clone(),
                "pending": [m._pending_batches for m in self._dual], "torch_rng": torch.cuda.get_rng_state(self.device)}
        side = self._stream = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up on the capture stream: cuDNN autotune, lazy module loads
                self._eager(st["images"], st["labels"], st["noise"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        with torch.no_grad():                        # the warm-up must not leak into the training state
            for (p, b), a in zip(snap["arena"], self.arenas):
                a.param.copy_(p); a.buf.copy_(b)
            for k, v in snap["model"].items():
                tensors[k].copy_(v)
            self.rng_offset.copy_(snap["rng"])
        for m, pend in zip(self._dual, snap["pending"]):
            m._pending_batches = pend
        torch.cuda.set_rng_state(snap["torch_rng"], self.device)
        pend0 = [m._pending_batches for m in self._dual]
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):
            st["loss"], st["parts"] = self._eager(st["images"], st["labels"], st["noise"])
        self._dual_per_iter = [m._pending_batches - p0 for m, p0 in zip(self._dual, pend0)]
        for m, p0 in zip(self._dual, pend0):
            m._pending_batches = p0                  # capture records launches, it does not run them

    def step(self, images: torch.Tensor, labels: torch.Tensor, noise: Optional[Dict[str, torch.Tensor]] = None):
        """One training iteration on device tensors; returns {'loss', 'losses' (l0..l3)} as DEVICE tensors."""
        self.model.train()
        if not self.use_graph:
            loss, parts = self._eager(images, labels, noise)
            self.iterations += 1
            return {"loss": loss, "losses": parts}
        if self._graph is None:
            self._capture(images, labels, noise)
        st = self._static
        st["images"].copy_(images, non_blocking=True)
        st["labels"].copy_(labels, non_blocking=True)
        if noise:
            for k, v in noise.items():
                st["noise"][k].copy_(v, non_blocking=True)
        self._graph.replay()
        for m, inc in zip(self._dual, self._dual_per_iter):
            m._pending_batches += inc
        self.iterations += 1
        return {"loss": st["loss"], "losses": st["parts"]}

    def close(self):
        """Drop the captured graph and its static buffers."""
        self._graph, self._static = None, {}
