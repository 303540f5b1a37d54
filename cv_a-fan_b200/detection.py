"""Drop-in for Detection/attack_algo.py's hot-path functions (Faster R-CNN flavour).


    PGD(x, image_batch, y, model, steps, eps, gamma, idx, randinit, clip)   (Detection/attack_algo.py:48-74)
    compute_loss(l1, l2, l3, l4)                                            (:21-27)
    mix_feature / get_sample_points                                         (:254-265 / :236-245)
The model contract is the reference's: model.train().forward({'x','adv','out_idx','flag'}, bb, lb) -> 4 losses.
"""
from .attack_algo import linfball_proj, l2ball_proj, pgd_loop  # noqa: F401
from .segmentation import get_sample_points, mix_feature, sat_sample_points  # noqa: F401  (identical maths in both reference files)


def compute_loss(loss1, loss2, loss3, loss4):
    return loss1.mean() + loss2.mean() + loss3.mean() + loss4.mean()


def PGD(x, image_batch, y=None, model=None, steps=3, eps=None, gamma=None, idx=1, randinit=False, clip=False,
        **extras):
    def tail_loss(x_adv):
        inputs = {"x": image_batch, "adv": x_adv, "out_idx": idx, "flag": "tail"}
        a_obj, a_trf, p_cls, p_trf = model.train().forward(inputs, y["bb"], y["lb"])
        return compute_loss(a_obj, a_trf, p_cls, p_trf)

    return pgd_loop(x, tail_loss, steps, gamma, eps, randinit, clip, **extras)


def rpn_roi_PGD(layer="roi", rpn_roi_output_dict=None, y=None, model=None, steps=1, eps=None, gamma=None,
                randinit=False, clip=False, only_roi_loss=True, **extras):
    """Detection/attack_algo.py:77-150: PGD on the pooled ROI feature (layer='roi') through
    model.train().forward({'adv': dict, 'out_idx': 'roi_tail', 'flag': 'clean'}, bb, lb).
    Reference defects handled explicitly: the 'roi' clip branch uses an undefined name (:111) --
    here it projects onto the eps-ball around the clean ROI feature; in the 'rpn' branch the update is commented out
    (:127-147), so the observable result is the (optionally randomly started) clean RPN feature as a leaf -- reproduced
    without the reference's wasted forward passes."""
    if layer == "roi":
        d = rpn_roi_output_dict
        anchor = d["roi_output_dict"]["roi_feature_map"].detach()

        def tail_loss(x_adv):
            d["roi_output_dict"]["roi_feature_map"] = x_adv
            a_obj, a_trf, p_cls, p_trf = model.train().forward({"adv": d, "out_idx": "roi_tail", "flag": "clean"},
                                                               y["bb"], y["lb"])
            if only_roi_loss:
                return p_cls.mean() + p_trf.mean()
            return compute_loss(a_obj, a_trf, p_cls, p_trf)

        d["roi_output_dict"]["roi_feature_map"] = pgd_loop(anchor, tail_loss, steps, gamma, eps, randinit, clip, **extras)
        return d
    if layer == "rpn":
        d = rpn_roi_output_dict
        anchor = d["rpn_feature_map_dict"]["rpn_feature"].detach()
        d["rpn_feature_map_dict"]["rpn_feature"] = pgd_loop(anchor, lambda x_adv: None, 0, gamma, eps, randinit, False,
                                                            **extras)
        return d
    raise AssertionError(f"unknown layer {layer!r}")


def adv_input(x=None, y=None, model=None, steps=3, eps=None, gamma=None, randinit=False, clip=False, **extras):
    """Detection/attack_algo.py:153-178: input-space PGD on the image batch, then clamp to [0, 1]."""
    import torch

    def tail_loss(x_adv):
        inputs = {"x": x_adv, "adv": None, "out_idx": -1, "flag": "clean"}
        return compute_loss(*model.train().forward(inputs, y["bb"], y["lb"]))

    return torch.clamp(pgd_loop(x, tail_loss, steps, gamma, eps, randinit, clip, **extras), 0, 1.0)


def nms(bboxes, scores, threshold):
    """Detection/support/layer/nms.py (`_C.nms`): indices of the kept boxes, ascending, like the reference returns them
    (`nms.cu:126-130`).  The suppression itself never leaves the GPU; only sizing the variable-length result
    synchronises (as torch.nonzero does).  `ops.nms_flags` gives the fixed-size device-side form."""
    import torch
    from . import ops
    if bboxes.numel() == 0:
        return torch.empty(0, dtype=torch.long, device=bboxes.device)
    keep, _ = ops.nms_flags(bboxes.float().contiguous(), scores.float().contiguous(), threshold)
    return torch.nonzero(keep).squeeze(1)


class _ROIAlign(__import__("torch").autograd.Function):
    """Detection/support/layer/roi_align.py:11-44 on the sm_100a kernels."""

    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        import torch
        from . import _lib
        oh, ow = (output_size, output_size) if isinstance(output_size, int) else output_size
        input, roi = input.contiguous(), roi.contiguous().float()
        n, c, h, w = input.shape
        out = torch.empty(roi.shape[0], c, oh, ow, dtype=input.dtype, device=input.device)
        _lib.check(_lib.lib().afan_roi_align_fwd_f32(_lib.f32(input, "input"), _lib.f32(roi, "roi"), _lib.f32(out), n, c, h, w,
                                                     roi.shape[0], oh, ow, float(spatial_scale), int(sampling_ratio),
                                                     _lib.stream()), "afan_roi_align_fwd_f32")
        ctx.save_for_backward(roi)
        ctx.cfg = (n, c, h, w, oh, ow, float(spatial_scale), int(sampling_ratio))
        return out

    @staticmethod
    def backward(ctx, grad_output):
        import torch
        from . import _lib
        (roi,) = ctx.saved_tensors
        n, c, h, w, oh, ow, scale, ratio = ctx.cfg
        grad_output = grad_output.contiguous()
        dfeat = torch.empty(n, c, h, w, dtype=grad_output.dtype, device=grad_output.device)
        _lib.check(_lib.lib().afan_roi_align_bwd_f32(_lib.f32(grad_output, "grad_output"), _lib.f32(roi), _lib.f32(dfeat), n, c,
                                                     h, w, roi.shape[0], oh, ow, scale, ratio, _lib.stream()),
                   "afan_roi_align_bwd_f32")
        return dfeat, None, None, None, None


roi_align = _ROIAlign.apply


class ROIAlign(__import__("torch").nn.Module):
    """Detection/support/layer/roi_align.py:50-68."""

    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size, self.spatial_scale, self.sampling_ratio = output_size, spatial_scale, sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)
