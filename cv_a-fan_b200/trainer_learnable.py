"""Learnable-eta multi-layer A-FAN: Classification/main_learnable.py:175-262 on the B200 path.

Reference iteration: for EACH of 9 perturbation layers -- a prefix forward (train mode, detached), a full PGD
(steps x [tail fwd + dgrad + update]) -- then 9 adversarial tail forwards on clean + w[i]*(adv - clean), one clean
full forward, loss = (CE_clean + mean_i CE_adv_i)/2 + l1_coef*||w||_1, two SGDs (network; w with its own lr, no weight
decay) and the sum-to-one projection of w (:369-378).  That is 9x the hot path per batch.  Here:

  * nested prefix cache   the 9 prefixes nest, so ONE pass over the head yields all 9 clean features; a BatchNorm
                          layer that lies in the prefixes of k points advances its running statistics k times
                          (`replay=k`), as the reference's k separate prefix forwards do.
  * batched ascent        (batched=True) all 9 ascents advance together: per step ONE pass over the tail in which
                          the adversarial feature of point i joins the batch at layer idx_i as a new statistic group
                          (dual-BN with up to 9 groups), one backward, 9 fused PGD-step launches.  Per-group statistics
                          make this the same maths as 9 separate passes; only the ORDER in which the running averages
                          absorb the per-pass statistics changes (declared deviation; batched=False keeps the
                          reference's pass order exactly).
  * batched final tails   the 9 adversarial tail forwards run as one progressive pass as well.
  * w                     mixed in with torch.lerp (autograd gives d/dw), updated by a 9-element momentum SGD and
                          projected with sum_project; the network uses the flat-arena fused SGD of AfanTrainer.
"""
from typing import Sequence

import torch

from . import attack_algo, ops
from ._lib import AfanError
from .dual_bn import DualBatchNorm2d
from .resnet_s import BasicBlock
from .trainer import AfanTrainer

DEFAULT_POINTS = (4, 8, 11, 14, 18, 21, 24, 28, 31)          # main_learnable.py:59 (ResNet-56, 34 layers)


def sum_project(w: torch.Tensor, K: int = 9) -> torch.Tensor:
    """main_learnable.py:369-378: shift w so that it sums to one."""
    return w - (w.sum(dim=0) - 1) / K


class LearnableEtaTrainer(AfanTrainer):
    def __init__(self, model, *, perturb_idx_list: Sequence[int] = DEFAULT_POINTS, w_lr: float = 0.01,
                 l1_coef: float = 1.0, batched: bool = True, steps: int = 3, gamma: float = 1.0, **kw):
        kw.pop("perturb_idx", None)
        kw.pop("head_cache", None)
        super().__init__(model, perturb_idx=perturb_idx_list[0], steps=steps, gamma=gamma, **kw)
        self.points = [int(k) for k in perturb_idx_list]
        if sorted(self.points) != self.points or len(set(self.points)) != len(self.points):
            raise AfanError("perturb_idx_list must be strictly increasing")
        if len(self.points) != model.w.numel():
            raise AfanError(f"model.w has {model.w.numel()} entries but {len(self.points)} perturbation layers were given")
        if self.norm != "linf":
            raise AfanError("the learnable-eta trainer implements the reference's L-inf ascent only")
        self.w_lr, self.l1_coef, self.batched = float(w_lr), float(l1_coef), bool(batched)
        self.w_buf = torch.zeros_like(model.w.data)
        self._pt_ws = None

    # the network parameters live in the arena; w has its own optimiser (main_learnable.py:78-88)
    def _build_arena(self, used):
        super()._build_arena([p for p in used if p is not self.model.w])
        self.model.w.grad = torch.zeros_like(self.model.w.data)

    def _tail_layers(self, x, start, end, groups):
        layers = self.model.sequential_model
        for li in range(start, end):
            m = layers[li]
            x = m(x, groups=groups) if isinstance(m, (BasicBlock, DualBatchNorm2d)) else m(x)
        return x

    def _progressive_tail(self, xs):
        """xs[i] enters at layer points[i] as statistic group i of a growing batch; returns [P*n, classes]."""
        pts, L = self.points, self.L
        x, g = None, 0
        for i, k in enumerate(pts):
            x = xs[i] if x is None else torch.cat([x, xs[i]], dim=0)
            g += 1
            x = self._tail_layers(x, k, pts[i + 1] if i + 1 < len(pts) else L, g)
        return x

    def _ascent_batched(self, anchors, target, norms):
        n, P = anchors[0].shape[0], len(anchors)
        xs = []
        for i, a in enumerate(anchors):
            x_adv = torch.empty_like(a)
            if self.randinit:
                ops.pgd_init(a, self.eps, seed=self.seed + i, offset_device=self.rng_offset, out=x_adv)
            else:
                x_adv.copy_(a)
            xs.append(x_adv.requires_grad_(True))
        if self.randinit:
            self.rng_offset += max((a.numel() + 3) // 4 for a in anchors)
        for t in range(self.steps):
            logits = self._progressive_tail(xs)
            loss = sum(self.criterion(logits[i * n:(i + 1) * n], target) for i in range(P))
            grads = torch.autograd.grad(loss, xs, only_inputs=True)
            last = t == self.steps - 1
            for i in range(P):
                ops.pgd_linf_step_(grads[i].contiguous(), anchors[i], xs[i].data, self.gamma, self.eps, self.clip,
                                   norms_out=norms[i] if last else None, workspace=self._pt_ws[i])
        return xs

    def _iteration(self, images, target, noise, norms_out, ws):
        model, pts, L, n = self.model, self.points, self.L, images.shape[0]
        P = len(pts)
        if self._pt_ws is None:
            self._pt_ws = [ops.norms_workspace(n, self.device) for _ in range(P)]
            self._pt_norms = torch.zeros(P, 2, n, dtype=torch.float32, device=self.device)
        # 1. nested prefix pass: all 9 clean features in one sweep over the head (main_learnable.py:202-204)
        anchors, x, prev = [], images, 0
        with torch.no_grad():
            for s, k in enumerate(pts):
                x = model(x, end_point=k, start_point=prev, replay=P - s)
                anchors.append(x.contiguous())
                prev = k
        # 2. the 9 ascents (:205-215)
        with self._params_frozen():
            if self.batched:
                xs = self._ascent_batched(anchors, target, self._pt_norms)
            else:
                xs = []
                for i, k in enumerate(pts):
                    extras = dict(norms_out=self._pt_norms[i], workspace=self._pt_ws[i])
                    if self.randinit:
                        extras.update(rng="philox", seed=self.seed + i, offset_device=self.rng_offset)
                    xs.append(attack_algo.PGD(anchors[i], self.criterion, y=target, model=model, steps=self.steps,
                                              gamma=self.gamma, start_idx=k, layer_number=L, eps=self.eps,
                                              randinit=self.randinit, clip=self.clip, **extras))
                if self.randinit:
                    self.rng_offset += max((a.numel() + 3) // 4 for a in anchors)
        # 3. clean + w[i] * (adv - clean) (:226) and the adversarial tails (:227)
        mixed = [torch.lerp(anchors[i], xs[i].detach(), model.w[i]) for i in range(P)]
        if self.batched:
            logits = self._progressive_tail(mixed)
            outs = [logits[i * n:(i + 1) * n] for i in range(P)]
        else:
            outs = [model(mixed[i], end_point=L, start_point=pts[i]) for i in range(P)]
        out_clean = model(images, end_point=L, start_point=0)                                   # :228
        loss_adv = sum(self.criterion(o, target) for o in outs)                                  # :238-240
        loss = (self.criterion(out_clean, target) + loss_adv / P) / 2 + model.w.abs().sum() * self.l1_coef   # :241-244
        norms_out.copy_(self._pt_norms[0])
        return loss, out_clean

    def _optimize(self, loss):
        w = self.model.w
        w.grad.zero_()
        super()._optimize(loss)                                   # network: arena zero, backward, all-reduce, fused SGD
        g = w.grad
        if self.world > 1:
            torch.distributed.all_reduce(g, group=self.pg)
            g = g / self.world
        self.w_buf.mul_(self.momentum).add_(g)                    # torch.optim.SGD(w, lr=w_lr, momentum, wd=0), :84-88
        w.data.add_(self.w_buf, alpha=-self.w_lr)
        w.data.copy_(sum_project(w.data, K=w.numel()))           # :252-253

    def _snapshot(self):
        s = super()._snapshot()
        s["w"], s["w_buf"] = self.model.w.data.clone(), self.w_buf.clone()
        return s

    def _restore(self, s):
        super()._restore(s)
        self.model.w.data.copy_(s["w"])
        self.w_buf.copy_(s["w_buf"])

    def point_norms(self):
        """[P, 2, N]: per perturbation layer, per-sample ||delta||_2 and ||delta||_inf of the last iteration (:219-223)."""
        return self._pt_norms
