"""The Detection flavour of the A-FAN training iteration: Detection/train_aug_final.py:78-163.

One reference iteration is

    #1  model({'flag': 'head', 'out_idx': se})                      backbone up to layer<se>        (:83-85)
    #2  model({'flag': 'clean', 'out_idx': 'roi_head'})             backbone + RPN + proposal NMS + ROI head  (:87-88)
    PGD (1 step) on the layer<se> feature: tail = rest of backbone + RPN + NMS + ROIAlign + ROI head + 4 losses  (:90-98)
    rpn_roi_PGD (1 step) on the pooled ROI feature: tail = two Linear layers + 2 losses                   (:100-110)
    mix_feature / uniform noise on the ROI-side adversarial feature                                      (:114-118)
    5 SAT points on the clean -> adversarial segment, masked mix_feature                                 (:120-129)
    six training forwards: clean, four tails from the SAT points, one ROI tail                           (:138-149)
    loss = (l0 + .. + l4) / 3 * (1 - w) + l5 / 3 * w,  SGD                                                (:159-163)

What this trainer does differently (same values on the same inputs and random draws):

  * head cache: forwards #1, #2 and the clean forward run the SAME frozen-BatchNorm backbone on the SAME images, and #2
    and the clean forward also the same RPN convolutions and the same proposal NMS -> ONE backbone sweep (with the
    autograd graph the clean loss needs), ONE RPN prediction and ONE NMS serve all three.  At `se = 3` (BASELINE
    config 4: the tail of the split is the identity) that removes two of three ResNet-101 C4 sweeps per iteration.
  * the ascent steps, the random start, the SAT points with their masked mix_feature (one launch for four points), NMS
    and ROIAlign are this package's sm_100a kernels; the per-image loss loops are segment reductions (faster_rcnn.py).
  * one flat SGD arena for the trainable parameters (one optimiser launch).

The random draws (candidate sampling, random start, noise) follow the reference order, so with `rng='reference'` and the
model's `sampler='reference'` an iteration seeded like the reference selects the same anchors and proposals.
"""
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import detection, ops, segmentation
from ._lib import AfanError
from .trainer_seg import _Arena


def warmup_multistep_lr(step: int, base_lr: float, milestones=(50000, 70000), gamma: float = 0.1, factor: float = 0.3333,
                        num_iters: int = 500) -> float:
    """Detection/extension/lr_scheduler.py:7-23 (WarmUpMultiStepLR) in closed form: the learning rate of iteration
    `step` (0-based: the value the optimiser uses for its step+1-th update)."""
    lr = base_lr * gamma ** sum(1 for m in milestones if step >= m)
    if step < num_iters:
        lr *= (1 - factor) * (step / num_iters) + factor
    return lr


class DetAfanTrainer:
    def __init__(self, model: nn.Module, *, pertub_idx_se: int = 3, pertub_idx_sd: str = "roi", steps: int = 1, eps: float = 2.0,
                 gamma_se: float = 0.5, gamma_sd: float = 0.1, randinit: bool = False, clip: bool = False, mix_layer: str = "0000",
                 noise_sd: float = 0.0, only_roi_sd: bool = False, mix_sd: bool = False, sd_adv_loss_weight: float = 0.5,
                 lr: float = 0.001, momentum: float = 0.9, weight_decay: float = 0.0005, head_cache: bool = True,
                 rng: str = "philox", seed: int = 0):
        """Flag names / units follow Detection/train_aug_final.py:199-211 (gamma_* in 1/255; the reference hard-codes
        eps = 2/255 and one ascent step at :93-95,104-106 -- the defaults here)."""
        if pertub_idx_se not in (1, 2, 3):
            raise AfanError("pertub_idx_se must be 1, 2 or 3 (backbone/resnet101_ori.py:205-235)")
        if pertub_idx_sd != "roi":
            # train_aug_final.py:88 indexes ['roi_output_dict'] unconditionally: 'rpn' raises KeyError in the reference
            raise AfanError("pertub_idx_sd must be 'roi' (the only value the reference iteration can run)")
        if len(mix_layer) != 4 or any(ch not in "01" for ch in mix_layer):
            raise AfanError("mix_layer is four 0/1 flags, e.g. '0101' (train_aug_final.py:75-76)")
        self.model, self.se, self.sd = model, pertub_idx_se, pertub_idx_sd
        self.steps, self.eps = int(steps), eps / 255.0
        self.gamma_se, self.gamma_sd, self.gamma_sd_raw = gamma_se / 255.0, gamma_sd / 255.0, float(gamma_sd)
        self.randinit, self.clip, self.mix_sd, self.noise_sd = randinit, clip, mix_sd, float(noise_sd)
        self.only_roi_sd, self.w_sd = only_roi_sd, float(sd_adv_loss_weight)
        self.mix = [int(ch) for ch in mix_layer]
        self.momentum, self.weight_decay = momentum, weight_decay
        self.head_cache, self.rng, self.seed = head_cache, rng, int(seed)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise AfanError("DetAfanTrainer needs the model on a CUDA device: there is no CPU path")
        # torch.optim.SGD skips parameters that never receive a gradient (frozen ones, and the backbone's unused `fc`)
        trainable = [p for n, p in model.named_parameters() if p.requires_grad and not n.startswith("features.fc.")]
        self.arena = _Arena(trainable, lr, self.device)
        self._lr = float(lr)
        self.rng_offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.iterations = 0

    def set_lr(self, lr: float):
        if lr != self._lr:
            self.arena.lr.fill_(lr)
            self._lr = float(lr)

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

    def _iteration(self, images, bboxes, labels, noise):
        model, se = self.model, self.se
        y = {"bb": bboxes, "lb": labels}
        rpn = None
        if self.head_cache:
            feat_k = model.features.head(images, se)                            # ONE sweep for #1, #2 and the clean forward
            features = model.features.tail(feat_k, se)
            feat_se = feat_k.detach()
            rpn = model.rpn_outputs(images, features)                           # shared by #2 and the clean forward
            sd_dict = model.roi_head_from_features(images, features, bboxes, labels, rpn)
        else:
            feat_se = model({"x": images, "adv": None, "out_idx": se, "flag": "head"}, bboxes, labels).detach()        # :83-85
            sd_dict = model({"x": images, "adv": None, "out_idx": "roi_head", "flag": "clean"}, bboxes, labels)       # :87
        clean_sd = sd_dict["roi_output_dict"]["roi_feature_map"].detach()                                           # :88

        adv_se = detection.PGD(feat_se, images, y=y, model=model, steps=self.steps, eps=self.eps, gamma=self.gamma_se, idx=se,
                               randinit=self.randinit, clip=self.clip, **self._extras(noise, "se", feat_se.numel()))   # :90-98
        adv_dict = detection.rpn_roi_PGD(layer=self.sd, rpn_roi_output_dict=sd_dict, y=y, model=model, steps=self.steps,
                                         eps=self.eps, gamma=self.gamma_sd, randinit=self.randinit, clip=self.clip,
                                         only_roi_loss=self.only_roi_sd, **self._extras(noise, "sd", clean_sd.numel()))  # :100-110
        adv_sd = adv_dict["roi_output_dict"]["roi_feature_map"].detach()                                            # :112
        if self.mix_sd:
            adv_sd = segmentation.mix_feature(clean_sd, adv_sd)                                                     # :114-115
        if self.noise_sd != 0:
            if noise is not None and "noise_sd" in noise:
                u = noise["noise_sd"].to(self.device)
            elif self.rng == "reference":
                u = torch.rand(adv_sd.shape).to(self.device)
            else:
                u = torch.rand(adv_sd.shape, device=self.device)
            adv_sd = adv_sd + (2.0 * u - 1.0) * self.gamma_sd_raw * self.noise_sd                                   # :116-117 (args.gamma_sd, not /255)
        adv_dict["roi_output_dict"]["roi_feature_map"] = adv_sd                                                     # :118
        pts = segmentation.sat_sample_points(feat_se, adv_se, 5, mix=self.mix)                                      # :120-129

        if self.head_cache:
            four = [model.losses_from_features(images, features, bboxes, labels, rpn)]                              # :138
        else:
            four = [model({"x": images, "adv": None, "out_idx": 0, "flag": "clean"}, bboxes, labels)]
        for i in range(1, 5):                                                                                       # :140-147
            four.append(model({"x": images, "adv": pts[i], "out_idx": se, "flag": "tail"}, bboxes, labels))
        four.append(model({"adv": adv_dict, "out_idx": "roi_tail", "flag": "clean"}, bboxes, labels))               # :148-149
        parts = torch.stack([detection.compute_loss(*f) for f in four])                                             # :151-156
        loss = (parts[:5].sum() / 3.0) * (1 - self.w_sd) + (parts[5] / 3.0) * self.w_sd                            # :159
        return loss, parts.detach()

    def step(self, images: torch.Tensor, bboxes: torch.Tensor, labels: torch.Tensor,
             noise: Optional[Dict[str, torch.Tensor]] = None):
        """One training iteration on device tensors (images [B,3,H,W] in [0,1], zero-padded gt boxes [B,G,4] and classes
        [B,G], as Detection/dataset/base.py's padding_collate_fn delivers them).  Returns {'loss', 'losses' (l0..l5)} as
        DEVICE tensors."""
        self.model.train()
        self.arena.grad.zero_()                                                                                     # :161
        loss, parts = self._iteration(images, bboxes, labels, noise)
        loss.backward()                                                                                             # :162
        ops.sgd_momentum_(self.arena.param, self.arena.grad, self.arena.buf, self.arena.lr, momentum=self.momentum,
                          weight_decay=self.weight_decay)                                                           # :163
        self.iterations += 1
        return {"loss": loss.detach(), "losses": parts}
