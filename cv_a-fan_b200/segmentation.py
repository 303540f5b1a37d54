"""Drop-in for Segmentation/attack_algo.py's hot-path functions (DeepLabv3+ flavour).

    PGD(x, image_batch, low_level_feat, criterion, y, model, steps, eps, gamma, idx, randinit, clip)
                                                                  (Segmentation/attack_algo.py:40-59)
    mix_feature(clean_feature, adv_feature)                       (:121-130)
    get_sample_points(pointx, pointy, number)                     (:108-118)
The model contract is the reference's dict API: model({'x','adv','out_idx','flag','low_level_feat'}).
"""
import torch

from . import ops
from .attack_algo import linfball_proj, l2ball_proj, pgd_loop  # noqa: F401  (same helpers as the reference file)
from ._lib import AfanError


def PGD(x, image_batch, low_level_feat, criterion, y=None, model=None, steps=3, eps=None, gamma=None, idx=1,
        randinit=False, clip=False, **extras):
    def tail_loss(x_adv):
        inputs = {"x": image_batch, "adv": x_adv, "out_idx": idx, "flag": "tail", "low_level_feat": low_level_feat}
        return criterion(model(inputs), y)

    return pgd_loop(x, tail_loss, steps, gamma, eps, randinit, clip, **extras)


def decoder_PGD(input_dict, image_batch, criterion, y=None, model=None, steps=3, eps=None, gamma=None, idx=1,
                randinit=False, clip=False, **extras):
    """Segmentation/attack_algo.py:61-84: PGD on the decoder-side feature input_dict['adv'] (ASPP / concat input);
    the tail is model({'x', 'adv': input_dict, 'out_idx': idx + '_tail', 'flag': 'clean'}).  Returns input_dict with
    'adv' replaced by the adversarial leaf.  The reference's clip branch references an undefined `x` (:81, NameError);
    the intended semantics -- projection onto the eps-ball around the CLEAN decoder feature -- is what runs here."""
    anchor = input_dict["adv"].detach()

    def tail_loss(x_adv):
        input_dict["adv"] = x_adv
        return criterion(model({"x": image_batch, "adv": input_dict, "out_idx": idx + "_tail", "flag": "clean"}), y)

    input_dict["adv"] = pgd_loop(anchor, tail_loss, steps, gamma, eps, randinit, clip, **extras)
    return input_dict


def adv_input(x=None, criterion=None, y=None, model=None, steps=3, eps=None, gamma=None, randinit=False, clip=False,
              **extras):
    """Segmentation/attack_algo.py:86-105: input-space PGD followed by clamp to [0, 1]."""
    def tail_loss(x_adv):
        return criterion(model({"x": x_adv, "adv": None, "out_idx": 0, "flag": "clean", "low_level_feat": None}), y)

    return torch.clamp(pgd_loop(x, tail_loss, steps, gamma, eps, randinit, clip, **extras), 0, 1.0)


def mix_feature(clean_feature, adv_feature):
    """Channel-dim mean/std of clean swapped for those of adv, one fused kernel (forward only: the
    reference applies it to detached / lerped features, main_aug_final.py:200-210)."""
    if clean_feature.requires_grad or adv_feature.requires_grad:
        clean_feature, adv_feature = clean_feature.detach(), adv_feature.detach()
    if not clean_feature.is_cuda:
        raise AfanError("mix_feature needs CUDA tensors: afan_b200 has no CPU path")
    return ops.mix_feature(clean_feature.contiguous(), adv_feature.contiguous())


def sat_sample_points(pointx, pointy, number, mix=None):
    """get_sample_points followed by the reference's per-point `adv_list[i] = mix_feature(clean, adv_list[i])`
    (main_aug_final.py:206-210 / train_aug_final.py:117-126) in ONE fused launch.  mix[i-1] selects point i
    (i = 1 .. number-1; the last point is pointy itself).  Returns [pointx, p_1', ..., p_{number-1}']."""
    m = number - 1
    mix = [False] * m if mix is None else [bool(f) for f in mix]
    if len(mix) != m:
        raise AfanError(f"mix must have {m} flags (points 1..{m})")
    if pointx.requires_grad or pointy.requires_grad:
        pointx, pointy = pointx.detach(), pointy.detach()
    percent = 1.0 / (number - 1)
    weights = [i * percent for i in range(1, number - 1)] + [1.0]
    if not mix[-1]:                                   # un-mixed last point is pointy itself (no copy), like the reference
        weights, flags = weights[:-1], mix[:-1]
    else:
        flags = mix
    pts = []
    for s in range(0, len(weights), 4):                # 4 points per launch
        pts += ops.sat_mix(pointx.contiguous(), pointy.contiguous(), weights[s:s + 4], flags[s:s + 4])
    if not mix[-1]:
        pts.append(pointy)
    return [pointx] + pts


def get_sample_points(pointx, pointy, number):
    """[x, lerp(x, y, i/(number-1)) for i in 1..number-2, y] (SAT points on the clean->adv segment)."""
    percent = 1.0 / (number - 1)
    pts = [pointx]
    for i in range(1, number - 1):
        pts.append(torch.lerp(pointx, pointy, i * percent))
    pts.append(pointy)
    return pts
