"""Splittable Faster R-CNN (ResNet-101 C4) for the Detection flavour of A-FAN.

Replaces, behind the reference's dict protocol and with a state dict whose keys are a subset of the reference's,

    Detection/model.py:18-185            Model.forward({'x','adv','out_idx','flag'}, gt_bboxes, gt_classes)
    Detection/model.py:229-436           Model.Detection (ROI head: labels, sampling, pooling, layer4, two Linear, losses)
    Detection/backbone/resnet101_ori.py:203-265   head / tail / clean split of the ResNet at out_idx in {1, 2, 3}
    Detection/rpn/region_proposal_network.py      anchors, RPN labels + sampling + losses, proposals (NMS)
    Detection/bbox.py, roi/pooler.py (ALIGN mode), extension/functional.py

What is different from the reference (B200-first, results identical on the same inputs and the same random draws):

  * NMS and ROIAlign run on the sm_100a kernels of this package (`detection.nms`, `detection.roi_align`): the greedy
    sweep never leaves the GPU (the reference copies the mask matrix to the host per image per forward, nms.cu:99-123).
  * every per-image Python loop is gone: the per-image mean losses (model.py:365-377, region_proposal_network.py:176-190)
    are one segment reduction (a [B, S] one-hot product: deterministic, no atomics), proposals of the whole batch are
    decoded / clipped / sorted in one pass, anchors and their inside-the-image index are built once per image size and
    cached on the device.
  * label assignment costs ONE host synchronisation per sampler (the two candidate counts), against five
    `nonzero()` synchronisations in the reference; the candidate lists are then built with `nonzero_static`.
  * `rpn_outputs()` / `roi_head()` expose the deterministic halves of a forward so that the training step
    (trainer_det.py) can run the backbone, the RPN convolutions and the proposal NMS ONCE per iteration for the three
    reference forwards that repeat them on identical inputs (train_aug_final.py:83-88,135-136).

Random sampling.  The reference draws `torch.randperm(len(candidates))` from the CPU default generator at six sites per
forward (three in the RPN, three in the ROI head).  `Sampler("reference")` makes the same draws in the same order, so a
run seeded like the reference selects the same anchors / proposals; `Sampler("device")` draws on the GPU.

BatchNorm is frozen: every BatchNorm2d of the backbone and of layer4 stays in eval mode with requires_grad False
(model.py:27-35,47-48), conv1 / bn1 / layer1 are frozen as well (backbone/resnet101.py:29-31).  Declared deviation: the
reference's 'rpn_tail' branch (model.py:98-113) is the one branch that does not re-freeze BatchNorm after the caller's
`model.train()`, so layer4 normalises with batch statistics there; that branch is unreachable from the training iteration
(train_aug_final.py:88 raises KeyError for pertub_idx_sd='rpn'), and here BatchNorm is frozen in every branch
(tests/test_oracle_vs_reference.py compares it with the reference's BatchNorm frozen by hand).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import detection, ops
from .dual_bn import frozen_bn_act
from .resnet_s import NormalizeByChannelMeanStd

IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


# ---------------------------------------------------------------------------------------------------
# box arithmetic (Detection/bbox.py), written on (x0, y0, x1, y1) columns without the stack / repeat temporaries
# ---------------------------------------------------------------------------------------------------
def box_deltas(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """bbox.py:42-52 calc_transformer: (dx/w, dy/h, log dw, log dh) taking src boxes to dst boxes."""
    sw, sh = src[..., 2] - src[..., 0], src[..., 3] - src[..., 1]
    dw, dh = dst[..., 2] - dst[..., 0], dst[..., 3] - dst[..., 1]
    scx, scy = (src[..., 0] + src[..., 2]) / 2, (src[..., 1] + src[..., 3]) / 2
    dcx, dcy = (dst[..., 0] + dst[..., 2]) / 2, (dst[..., 1] + dst[..., 3]) / 2
    return torch.stack(((dcx - scx) / sw, (dcy - scy) / sh, torch.log(dw / sw), torch.log(dh / sh)), dim=-1)


def apply_deltas(src: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """bbox.py:54-64 apply_transformer."""
    sw, sh = src[..., 2] - src[..., 0], src[..., 3] - src[..., 1]
    cx = t[..., 0] * sw + (src[..., 0] + src[..., 2]) / 2
    cy = t[..., 1] * sh + (src[..., 1] + src[..., 3]) / 2
    w, h = torch.exp(t[..., 2]) * sw, torch.exp(t[..., 3]) * sh
    return torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), dim=-1)


def pairwise_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bbox.py:66-82: IoU of a [B, P, 4] against b [B, G, 4] -> [B, P, G], broadcast instead of repeat().  A zero-area
    pair gives 0/0 = NaN exactly as in the reference (every comparison on it is false: the label stays -1)."""
    a, b = a.unsqueeze(-2), b.unsqueeze(-3)
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    w = (torch.min(a[..., 2], b[..., 2]) - torch.max(a[..., 0], b[..., 0])).clamp(min=0)
    h = (torch.min(a[..., 3], b[..., 3]) - torch.max(a[..., 1], b[..., 1])).clamp(min=0)
    inter = w * h
    return inter / (area_a + area_b - inter)


def clip_boxes(boxes: torch.Tensor, width: float, height: float) -> torch.Tensor:
    """bbox.py:89-92 with left = top = 0."""
    x = boxes[..., 0::2].clamp(min=0, max=width)
    y = boxes[..., 1::2].clamp(min=0, max=height)
    return torch.stack((x[..., 0], y[..., 0], x[..., 1], y[..., 1]), dim=-1)


def smooth_l1(diff: torch.Tensor, beta: float) -> torch.Tensor:
    """extension/functional.py:6-10, elementwise part."""
    diff = diff.abs()
    return torch.where(diff < beta, 0.5 * diff ** 2 / beta, diff - 0.5 * beta)


def per_image_losses(logits, targets, pred_deltas, gt_deltas, batch_indices, batch_size: int, beta: float):
    """The reference's per-image loops (model.py:365-377 / region_proposal_network.py:176-190) as segment reductions:
    cross entropy averaged over each image's samples; smooth-L1 summed over each image's foreground samples and divided
    by (4 * n_foreground + 1e-8).  An image without samples gives NaN (mean of nothing), as in the reference."""
    seg = F.one_hot(batch_indices, batch_size).to(logits.dtype).t()              # [B, S]
    ce = F.cross_entropy(logits, targets, reduction="none")
    cnt = seg.sum(dim=1)
    # an image without samples: NaN like the reference's mean of nothing, but without 0/0 in the backward of the others
    cross_entropies = torch.where(cnt > 0, (seg @ ce) / cnt.clamp(min=1.0), torch.full_like(cnt, float("nan")))
    fg = targets != 0
    # background rows may hold inf / NaN regression targets (log of a zero-area ratio): their difference is replaced by 0
    # BEFORE the loss, so neither the value nor the gradient sees them (the reference gathers the foreground rows)
    diff = torch.where(fg.unsqueeze(1), pred_deltas - gt_deltas, torch.zeros_like(pred_deltas))
    fg = fg.to(logits.dtype)
    smooth_l1_losses = (seg @ smooth_l1(diff, beta).sum(dim=1)) / (4.0 * (seg @ fg) + 1e-8)
    return cross_entropies, smooth_l1_losses


# ---------------------------------------------------------------------------------------------------
# random candidate selection
# ---------------------------------------------------------------------------------------------------
class Sampler:
    """`perm(n)` = a random permutation of n as a LongTensor on `device`.
    mode 'reference': drawn from the CPU default generator, like the reference's bare `torch.randperm(n)`;
    mode 'device': drawn by the GPU generator (no host-to-device copy)."""

    def __init__(self, mode: str = "reference"):
        if mode not in ("reference", "device"):
            raise ValueError(f"unknown sampler mode {mode!r}")
        self.mode = mode

    def perm(self, n: int, device) -> torch.Tensor:
        if self.mode == "reference":
            return torch.randperm(n).to(device, non_blocking=True)
        return torch.randperm(n, device=device)


def select_samples(labels: torch.Tensor, fg_cap: int, total: int, sampler: Sampler):
    """The selection both stages share (region_proposal_network.py:86-91, model.py:257-262): up to `fg_cap` random
    foreground candidates (label > 0), filled to `total` with random background candidates (label == 0), shuffled.
    Returns (batch index [S], candidate index [S]).  One host synchronisation (the two counts)."""
    fg_mask, bg_mask = labels > 0, labels == 0
    n_fg, n_bg = torch.stack((fg_mask.sum(), bg_mask.sum())).tolist()
    fg = torch.nonzero_static(fg_mask, size=n_fg)                                # row-major, like nonzero()
    bg = torch.nonzero_static(bg_mask, size=n_bg)
    fg = fg[sampler.perm(n_fg, labels.device)[:min(n_fg, fg_cap)]]
    bg = bg[sampler.perm(n_bg, labels.device)[:max(total - fg.shape[0], 0)]]
    sel = torch.cat((fg, bg), dim=0)
    sel = sel[sampler.perm(sel.shape[0], labels.device)]
    return sel[:, 0], sel[:, 1]


# ---------------------------------------------------------------------------------------------------
# backbone (torchvision-layout ResNet-101 with the reference's head / tail split)
# ---------------------------------------------------------------------------------------------------
class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, width: int = None, downsample: bool = False):
        super().__init__()
        width = planes if width is None else width
        self.conv1 = nn.Conv2d(inplanes, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))

    fused_bn = False            # set by FasterRCNN(fuse_frozen_bn=True): frozen BatchNorm + residual + ReLU as one launch

    def forward(self, x):
        if self.fused_bn and x.is_cuda:
            y = frozen_bn_act(self.conv1(x), self.bn1, relu=True)
            y = frozen_bn_act(self.conv2(y), self.bn2, relu=True)
            idn = x if self.downsample is None else frozen_bn_act(self.downsample[0](x), self.downsample[1])
            return frozen_bn_act(self.conv3(y), self.bn3, residual=idn, relu=True)
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


def _stage(inplanes: int, planes: int, depth: int, stride: int, base_width: int) -> nn.Sequential:
    width = int(planes * (base_width / 64.0))
    blocks = [Bottleneck(inplanes, planes, stride, width, downsample=(stride != 1 or inplanes != planes * 4))]
    blocks += [Bottleneck(planes * 4, planes, 1, width) for _ in range(depth - 1)]
    return nn.Sequential(*blocks)


class SplitResNet(nn.Module):
    """backbone/resnet101_ori.py:120-265.  `base_width` is the reference's `width_per_group` (64 = ResNet-101; the
    tests use a narrow one: the stage OUTPUT widths 256 / 512 / 1024 / 2048 do not depend on it)."""

    def __init__(self, layers=(3, 4, 23, 3), base_width: int = 64, num_classes: int = 1000):
        super().__init__()
        self.normal = NormalizeByChannelMeanStd(IMAGENET_MEAN, IMAGENET_STD)
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _stage(64, 64, layers[0], 1, base_width)
        self.layer2 = _stage(256, 128, layers[1], 2, base_width)
        self.layer3 = _stage(512, 256, layers[2], 2, base_width)
        self.layer4 = _stage(1024, 512, layers[3], 2, base_width)       # the ROI head's `hidden` (same module object)
        self.fc = nn.Linear(2048, num_classes)                          # never run; kept for checkpoint compatibility
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    fused_bn = False

    def stem(self, x):
        if self.fused_bn and x.is_cuda:
            return self.layer1(self.maxpool(frozen_bn_act(self.conv1(self.normal(x)), self.bn1, relu=True)))
        return self.layer1(self.maxpool(F.relu(self.bn1(self.conv1(self.normal(x))))))

    def head(self, x, out_idx: int):
        """resnet101_ori.py:205-235: image -> output of layer<out_idx>."""
        if out_idx not in (1, 2, 3):
            raise AssertionError(f"out_idx must be 1, 2 or 3, got {out_idx!r}")
        x = self.stem(x)
        for stage in (self.layer2, self.layer3)[:out_idx - 1]:
            x = stage(x)
        return x

    def tail(self, feature, out_idx: int):
        """resnet101_ori.py:237-249: output of layer<out_idx> -> layer3 output (identity for out_idx = 3)."""
        if out_idx not in (1, 2, 3):
            raise AssertionError(f"out_idx must be 1, 2 or 3, got {out_idx!r}")
        for stage in (self.layer2, self.layer3)[out_idx - 1:]:
            feature = stage(feature)
        return feature

    def forward(self, input_dict):
        flag = input_dict["flag"]
        if flag == "head":
            return self.head(input_dict["x"], input_dict["out_idx"])
        if flag == "tail":
            return self.tail(input_dict["adv"], input_dict["out_idx"])
        if flag == "clean":
            return self.tail(self.stem(input_dict["x"]), 1)
        raise AssertionError(f"unknown flag {flag!r}")


# ---------------------------------------------------------------------------------------------------
# region proposal network
# ---------------------------------------------------------------------------------------------------
class RegionProposalNetwork(nn.Module):
    def __init__(self, num_features_out, anchor_ratios, anchor_sizes, pre_nms_top_n, post_nms_top_n, anchor_smooth_l1_loss_beta,
                 sampler: Sampler):
        super().__init__()
        self._features = nn.Sequential(nn.Conv2d(num_features_out, 512, 3, padding=1), nn.ReLU())
        num_anchors = len(anchor_ratios) * len(anchor_sizes)
        self._anchor_objectness = nn.Conv2d(512, num_anchors * 2, 1)
        self._anchor_transformer = nn.Conv2d(512, num_anchors * 4, 1)
        self._anchor_ratios, self._anchor_sizes = list(anchor_ratios), list(anchor_sizes)
        self._pre_nms_top_n, self._post_nms_top_n = pre_nms_top_n, post_nms_top_n
        self._beta = anchor_smooth_l1_loss_beta
        self._sampler = sampler
        self._anchor_cache = {}

    # -- anchors ------------------------------------------------------------------------------------
    def generate_anchors(self, image_width, image_height, num_x_anchors, num_y_anchors) -> torch.Tensor:
        """region_proposal_network.py:194-225: [A, 4] float32 on the CPU, centre y major, then x, ratio, size; computed
        in float64 and rounded once, like the reference's numpy code."""
        ys = np.linspace(0, image_height, num_y_anchors + 2)[1:-1]
        xs = np.linspace(0, image_width, num_x_anchors + 2)[1:-1]
        ratios = np.array([a / b for a, b in self._anchor_ratios], dtype=np.float64)
        sizes = np.array(self._anchor_sizes, dtype=np.float64)
        cy, cx, r, s = (g.reshape(-1) for g in np.meshgrid(ys, xs, ratios, sizes, indexing="ij"))
        centre = torch.from_numpy(np.stack((cx, cy, s * np.sqrt(1 / r), s * np.sqrt(r)), axis=1)).float()
        half_w, half_h = centre[:, 2] / 2, centre[:, 3] / 2
        return torch.stack((centre[:, 0] - half_w, centre[:, 1] - half_h, centre[:, 0] + half_w, centre[:, 1] + half_h), dim=1)

    def anchors_for(self, image_width, image_height, fw, fh, device):
        """(anchors [A, 4], index of the anchors inside the image [A_in]) on `device`, cached per geometry."""
        key = (image_width, image_height, fw, fh, str(device))
        if key not in self._anchor_cache:
            a = self.generate_anchors(image_width, image_height, fw, fh)
            inside = (a[:, 0] >= 0) & (a[:, 1] >= 0) & (a[:, 2] <= image_width) & (a[:, 3] <= image_height)    # bbox.py:84-87
            self._anchor_cache[key] = (a.to(device), torch.nonzero(inside).squeeze(1).to(device))
        return self._anchor_cache[key]

    # -- the deterministic part -------------------------------------------------------------------
    def conv_feature(self, features):
        return self._features(features)

    def predict(self, rpn_feature):
        """objectness [B, A, 2], transformers [B, A, 4] in (y, x, anchor) order (region_proposal_network.py:53-57)."""
        b = rpn_feature.shape[0]
        obj = self._anchor_objectness(rpn_feature).permute(0, 2, 3, 1).reshape(b, -1, 2)
        trf = self._anchor_transformer(rpn_feature).permute(0, 2, 3, 1).reshape(b, -1, 4)
        return obj, trf

    # -- training losses ------------------------------------------------------------------------------
    def losses(self, objectnesses, transformers, anchors, inside, gt_bboxes_batch):
        """region_proposal_network.py:62-103."""
        b = objectnesses.shape[0]
        in_anchors = anchors[inside].unsqueeze(0).expand(b, -1, -1)
        in_obj, in_trf = objectnesses[:, inside], transformers[:, inside]
        ious = pairwise_iou(in_anchors, gt_bboxes_batch)                          # [B, A_in, G]
        anchor_max, anchor_assign = ious.max(dim=2)
        gt_max = ious.max(dim=1)[0]
        best_for_some_gt = ((ious > 0) & (ious == gt_max.unsqueeze(1))).any(dim=2)
        labels = torch.full_like(anchor_assign, -1)
        labels[anchor_max < 0.3] = 0
        labels[best_for_some_gt] = 1
        labels[anchor_max >= 0.7] = 1
        bi, ai = select_samples(labels, 128 * b, 256 * b, self._sampler)
        sel_anchors = in_anchors[bi, ai]
        gt_deltas = box_deltas(sel_anchors, gt_bboxes_batch[bi, anchor_assign[bi, ai]])
        return per_image_losses(in_obj[bi, ai], labels[bi, ai], in_trf[bi, ai], gt_deltas, bi, b, self._beta)

    # -- proposals --------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate_proposals(self, anchors, objectnesses, transformers, image_width, image_height) -> torch.Tensor:
        """region_proposal_network.py:227-271: decode, clip, rank by objectness, NMS at 0.7, keep the best
        post_nms_top_n, zero-pad to the longest list of the batch.  Ranking uses the foreground logit: the reference
        ranks by a softmax ACROSS ANCHORS of that logit (:247), a monotone map, so the order is the same."""
        boxes = clip_boxes(apply_deltas(anchors.unsqueeze(0), transformers.detach()), image_width, image_height)
        order = torch.sort(objectnesses.detach()[:, :, 1], dim=1, descending=True, stable=True)[1][:, :self._pre_nms_top_n]
        ranked = torch.gather(boxes, 1, order.unsqueeze(2).expand(-1, -1, 4)).contiguous()
        # one launch pair for the batch: a CTA per image sweeps on the device and stops at post_nms_top_n kept boxes
        kept, counts, _ = ops.nms_batched(ranked, 0.7, self._post_nms_top_n)
        return kept[:, :int(counts.max())].contiguous()                              # ONE synchronisation (the padded length)


# ---------------------------------------------------------------------------------------------------
# ROI head
# ---------------------------------------------------------------------------------------------------
class DetectionHead(nn.Module):
    """model.py:229-436 `Model.Detection`."""

    def __init__(self, hidden: nn.Module, num_hidden_out: int, num_classes: int, proposal_smooth_l1_loss_beta, sampler: Sampler):
        super().__init__()
        self.hidden = hidden
        self.num_classes = num_classes
        self._proposal_class = nn.Linear(num_hidden_out, num_classes)
        self._proposal_transformer = nn.Linear(num_hidden_out, num_classes * 4)
        self._beta = proposal_smooth_l1_loss_beta
        self._sampler = sampler
        self.register_buffer("_norm_mean", torch.tensor([0.0, 0.0, 0.0, 0.0]), persistent=False)
        self.register_buffer("_norm_std", torch.tensor([0.1, 0.1, 0.2, 0.2]), persistent=False)

    def pool(self, features, boxes, batch_indices):
        """roi/pooler.py:35-43 (ALIGN): ROIAlign 14x14 at scale 1/16, adaptive sampling, then 2x2 max pooling."""
        rois = torch.cat((batch_indices.view(-1, 1).to(boxes.dtype), boxes), dim=1)
        return F.max_pool2d(detection.roi_align(features, rois, (14, 14), 1 / 16, 0), 2, 2)

    def embed(self, pooled):
        # adaptive_max_pool2d(., 1) of model.py:276 as a plain reduction (the pooling kernel and its atomic backward cost
        # 18 ms per iteration at config 4); on tied maxima -- zeros after the block's ReLU -- the gradient is shared instead
        # of given to the first, and the ReLU backward right below it zeroes both
        return self.hidden(pooled).amax(dim=(2, 3), keepdim=True)                      # [S, 2048, 1, 1]

    def assign(self, proposal_bboxes, gt_classes_batch, gt_bboxes_batch):
        """model.py:246-268: label every proposal, select 128 per image (<= 32 foreground), build the targets."""
        b = proposal_bboxes.shape[0]
        ious = pairwise_iou(proposal_bboxes, gt_bboxes_batch)
        max_iou, assign = ious.max(dim=2)
        labels = torch.full_like(assign, -1)
        labels[max_iou < 0.5] = 0
        fg = max_iou >= 0.5
        labels = torch.where(fg, gt_classes_batch.gather(1, assign), labels)
        bi, pi = select_samples(labels, 32 * b, 128 * b, self._sampler)
        boxes = proposal_bboxes[bi, pi]
        gt_deltas = box_deltas(boxes, gt_bboxes_batch[bi, assign[bi, pi]])
        return boxes, bi, labels[bi, pi], gt_deltas

    def classify(self, roi_feature_map):
        hidden = roi_feature_map.reshape(roi_feature_map.shape[0], self._proposal_class.in_features)
        return self._proposal_class(hidden), self._proposal_transformer(hidden)

    def losses(self, proposal_classes, proposal_transformers, gt_classes, gt_deltas, batch_size, batch_indices):
        """model.py:355-379."""
        own = proposal_transformers.view(-1, self.num_classes, 4)[torch.arange(gt_classes.shape[0], device=gt_classes.device), gt_classes]
        gt = (gt_deltas - self._norm_mean) / self._norm_std
        return per_image_losses(proposal_classes, gt_classes, own, gt, batch_indices, batch_size, self._beta)

    def head(self, features, proposal_bboxes, gt_classes_batch, gt_bboxes_batch) -> dict:
        """return_type='head' (model.py:283-320): everything up to the pooled 2048-vector the ROI-side PGD perturbs."""
        boxes, bi, gt_classes, gt_deltas = self.assign(proposal_bboxes, gt_classes_batch, gt_bboxes_batch)
        return {"roi_feature_map": self.embed(self.pool(features, boxes, bi)), "gt_proposal_classes": gt_classes,
                "gt_proposal_transformers": gt_deltas,
                "batch_size": torch.tensor([[features.shape[0]]]), "batch_indices": bi}      # host tensor: no .item() sync later

    def tail(self, d: dict):
        """return_type='tail' (model.py:322-336)."""
        classes, transformers = self.classify(d["roi_feature_map"])
        ce, l1 = self.losses(classes, transformers, d["gt_proposal_classes"], d["gt_proposal_transformers"],
                             int(d["batch_size"][0]), d["batch_indices"])
        return classes, transformers, ce, l1

    def forward_train(self, features, proposal_bboxes, gt_classes_batch, gt_bboxes_batch):
        return self.tail(self.head(features, proposal_bboxes, gt_classes_batch, gt_bboxes_batch))

    def forward_eval(self, features, proposal_bboxes):
        b, p = proposal_bboxes.shape[:2]
        bi = torch.arange(b, device=features.device).repeat_interleave(p)
        classes, transformers = self.classify(self.embed(self.pool(features, proposal_bboxes.reshape(-1, 4), bi)))
        return classes.view(b, p, -1), transformers.view(b, p, -1)

    @torch.no_grad()
    def generate_detections(self, proposal_bboxes, proposal_classes, proposal_transformers, image_width, image_height):
        """model.py:381-416: per image and foreground class, NMS at 0.3 over the decoded boxes."""
        b = proposal_bboxes.shape[0]
        t = proposal_transformers.view(b, -1, self.num_classes, 4) * self._norm_std + self._norm_mean
        boxes = clip_boxes(apply_deltas(proposal_bboxes.unsqueeze(2), t), image_width, image_height)
        probs = F.softmax(proposal_classes, dim=-1)
        out = [[], [], [], []]
        for i in range(b):
            for c in range(1, self.num_classes):
                kept = detection.nms(boxes[i, :, c].contiguous(), probs[i, :, c].contiguous(), 0.3)
                out[0].append(boxes[i, :, c][kept])
                out[1].append(torch.full((kept.shape[0],), c, dtype=torch.int))
                out[2].append(probs[i, :, c][kept])
                out[3].append(torch.full((kept.shape[0],), i, dtype=torch.long))
        return tuple(torch.cat(o, dim=0) for o in out)


# ---------------------------------------------------------------------------------------------------
# the model
# ---------------------------------------------------------------------------------------------------
class FasterRCNN(nn.Module):
    """Detection/model.py `Model` with the ResNet-101 backbone of backbone/resnet101.py (ALIGN pooler)."""

    def __init__(self, num_classes: int, anchor_ratios=((1, 2), (1, 1), (2, 1)), anchor_sizes=(128, 256, 512),
                 rpn_pre_nms_top_n: int = 12000, rpn_post_nms_top_n: int = 2000, anchor_smooth_l1_loss_beta: float = 1.0,
                 proposal_smooth_l1_loss_beta: float = 1.0, layers=(3, 4, 23, 3), base_width: int = 64, sampler: str = "reference",
                 fuse_frozen_bn: bool = True):
        """fuse_frozen_bn: every frozen BatchNorm (+ residual add + ReLU) of the backbone and of layer4 runs as ONE
        hand-written launch per direction (csrc/afan_bn.cu: afan_bn_affine_f32 / _bwd_f32) from a cached (scale, shift)
        table, instead of the library's eval BatchNorm + add + ReLU kernels; False keeps the library sequence."""
        super().__init__()
        self.sampler = Sampler(sampler)
        self.features = SplitResNet(layers, base_width)
        self.rpn = RegionProposalNetwork(1024, anchor_ratios, anchor_sizes, rpn_pre_nms_top_n, rpn_post_nms_top_n,
                                         anchor_smooth_l1_loss_beta, self.sampler)
        self.detection = DetectionHead(self.features.layer4, 2048, num_classes, proposal_smooth_l1_loss_beta, self.sampler)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.requires_grad = m.bias.requires_grad = False
        for m in (self.features.conv1, self.features.layer1):
            for p in m.parameters():
                p.requires_grad = False
        for m in self.modules():
            if isinstance(m, (Bottleneck, SplitResNet)):
                m.fused_bn = bool(fuse_frozen_bn)

    def train(self, mode: bool = True):
        # the reference calls model.train() before EVERY forward (train_aug_final.py:86-158, attack_algo.py:59); walking ~400
        # modules ten times per iteration is host time for nothing when the mode does not change
        mode = bool(mode)
        if getattr(self, "_mode_set", None) == mode and all(m.training == mode for m in (self, self.features, self.rpn, self.detection)):
            return self
        super().train(mode)
        for m in self.modules():                       # frozen statistics in every mode (model.py:47-48)
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        self._mode_set = mode
        return self

    # -- the halves of a training forward, exposed for the head cache of trainer_det -------------------
    def geometry(self, image, features):
        h, w = image.shape[2], image.shape[3]
        return self.rpn.anchors_for(w, h, features.shape[3], features.shape[2], features.device) + (w, h)

    def rpn_outputs(self, image, features, rpn_feature=None):
        """Deterministic: RPN convolution, predictions, proposals."""
        anchors, inside, w, h = self.geometry(image, features)
        rpn_feature = self.rpn.conv_feature(features) if rpn_feature is None else rpn_feature
        obj, trf = self.rpn.predict(rpn_feature)
        proposals = self.rpn.generate_proposals(anchors, obj, trf, w, h)
        return {"objectnesses": obj, "transformers": trf, "proposals": proposals, "anchors": anchors, "inside": inside}

    def rpn_losses(self, r: dict, gt_bboxes_batch):
        return self.rpn.losses(r["objectnesses"], r["transformers"], r["anchors"], r["inside"], gt_bboxes_batch)

    def losses_from_features(self, image, features, gt_bboxes_batch, gt_classes_batch, r: dict = None):
        """model.py:58-75 after `features = self.features(...)`: RPN losses (3 draws), proposals, ROI losses (3 draws)."""
        r = self.rpn_outputs(image, features) if r is None else r
        a_obj, a_trf = self.rpn_losses(r, gt_bboxes_batch)
        _, _, p_cls, p_trf = self.detection.forward_train(features, r["proposals"], gt_classes_batch, gt_bboxes_batch)
        return a_obj, a_trf, p_cls, p_trf

    def roi_head_from_features(self, image, features, gt_bboxes_batch, gt_classes_batch, r: dict = None) -> dict:
        """model.py:117-139 after the backbone."""
        r = self.rpn_outputs(image, features) if r is None else r
        a_obj, a_trf = self.rpn_losses(r, gt_bboxes_batch)
        return {"anchor_objectness_losses": a_obj, "anchor_transformer_losses": a_trf,
                "roi_output_dict": self.detection.head(features, r["proposals"], gt_classes_batch, gt_bboxes_batch)}

    # -- the reference's protocol -------------------------------------------------------------------------
    def forward(self, input_dict: dict, gt_bboxes_batch=None, gt_classes_batch=None):
        flag, out_idx = input_dict["flag"], input_dict.get("out_idx")
        if flag == "head":
            return self.features(input_dict)
        if flag not in ("tail", "clean"):
            raise AssertionError(f"unknown flag {flag!r}")
        if not self.training:
            features = self.features(input_dict)
            image = input_dict["x"]
            r = self.rpn_outputs(image, features)
            classes, transformers = self.detection.forward_eval(features, r["proposals"])
            return self.detection.generate_detections(r["proposals"], classes, transformers, image.shape[3], image.shape[2])
        if isinstance(out_idx, int):
            return self.losses_from_features(input_dict["x"], self.features(input_dict), gt_bboxes_batch, gt_classes_batch)
        if out_idx == "roi_head":
            return self.roi_head_from_features(input_dict["x"], self.features(input_dict), gt_bboxes_batch, gt_classes_batch)
        if out_idx == "roi_tail":
            d = input_dict["adv"]
            _, _, p_cls, p_trf = self.detection.tail(d["roi_output_dict"])
            return d["anchor_objectness_losses"], d["anchor_transformer_losses"], p_cls, p_trf
        if out_idx == "rpn_head":
            image, features = input_dict["x"], self.features(input_dict)
            anchors, _, w, h = self.geometry(image, features)
            return {"features": features, "image_height": torch.tensor([[h]]), "image_width": torch.tensor([[w]]),
                    "anchor_bboxes": anchors.unsqueeze(0).expand(image.shape[0], -1, -1),
                    "rpn_feature_map_dict": {"batch_size": torch.tensor([[image.shape[0]]]),
                                             "rpn_feature": self.rpn.conv_feature(features)}}
        if out_idx == "rpn_tail":
            d = input_dict["adv"]
            features, w, h = d["features"], int(d["image_width"][0]), int(d["image_height"][0])
            anchors, inside = self.rpn.anchors_for(w, h, features.shape[3], features.shape[2], features.device)
            obj, trf = self.rpn.predict(d["rpn_feature_map_dict"]["rpn_feature"])
            r = {"objectnesses": obj, "transformers": trf, "anchors": anchors, "inside": inside,
                 "proposals": self.rpn.generate_proposals(anchors, obj, trf, w, h)}
            return self.losses_from_features(None, features, gt_bboxes_batch, gt_classes_batch, r)
        raise AssertionError(f"unknown out_idx {out_idx!r}")

    # -- checkpoints ------------------------------------------------------------------------------------
    def load_reference_state_dict(self, state_dict: dict) -> int:
        """model.py:200-211: copy every tensor whose key this model also has (the reference checkpoint additionally holds
        `_bn_modules.<i>.*` aliases of the BatchNorm tensors and `detection.hidden.*` aliases of `features.layer4.*`,
        which are the same storage here).  Returns the number of keys taken."""
        own = self.state_dict()
        take = {k: v for k, v in state_dict.items() if k in own}
        own.update(take)
        self.load_state_dict(own)
        return len(take)
