"""TEST INFRASTRUCTURE ONLY -- restatement of ONE iteration of the reference's Detection training loop
(Detection/train_aug_final.py:78-163; the loop body is inline in _train(), so it cannot be imported) around the UNMODIFIED
reference model (Detection/model.py, rpn/, roi/, bbox.py, backbone/resnet101_ori.py) and the UNMODIFIED
Detection/attack_algo.py functions, imported under oracle/ref_shim.py.

The reference model needs `support._C` (nms, roi_align_forward/backward), a CUDA extension that cannot be built here
(SURVEY F8).  The shim supplies `support.layer.nms` / `support.layer.roi_align` from oracle/afan_oracle.c, the C
restatement of Detection/support/src/cuda/nms.cu and ROIAlign_cuda.cu that tests/test_oracle_golden.py pins against the
reference's own NMS golden (9770 -> 1934 boxes).  Everything else is the reference's code, executed on the CPU.

`generate()` writes tests/golden/det_step.npz; the GPU test replays the same inputs and the same CPU random stream
through cv_a-fan_b200.trainer_det.DetAfanTrainer.  Only tests/ and this generator may import this file.
"""
import os
import sys
import types
import zlib

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden", "det_step.npz")

CASES = {
    "A": dict(se=3, gamma_se=0.5, gamma_sd=0.1, randinit=True, clip=False, mix_layer="0101", noise_sd=0.5, only_roi_sd=True,
              mix_sd=True, w=0.5),
    "B": dict(se=1, gamma_se=1.0, gamma_sd=0.5, randinit=False, clip=False, mix_layer="1000", noise_sd=0.0, only_roi_sd=False,
              mix_sd=False, w=0.3),
}
# a narrow, shallow ResNet of the reference's own class (stage OUTPUT widths stay 256 / 512 / 1024 / 2048)
MODEL = dict(num_classes=5, layers=(1, 1, 2, 1), base_width=8, anchor_ratios=[(1, 2), (1, 1), (2, 1)], anchor_sizes=[32, 64, 96],
             pre_nms=300, post_nms=60, beta=1.0)
BATCH, HEIGHT, WIDTH, ITERS, LR, MOMENTUM, WD = 2, 160, 192, 2, 0.002, 0.9, 0.0005
FULL = ("rpn._anchor_objectness.weight", "rpn._anchor_objectness.bias", "detection._proposal_class.weight",
        "detection._proposal_transformer.bias", "features.layer3.1.conv1.weight", "features.layer2.0.conv3.weight")


def procedural_init(model: nn.Module, seed: int):
    """Deterministic initialisation keyed on the state-dict KEY (not on construction or key order): the reference model
    and cv_a-fan_b200.faster_rcnn hold the same tensors under the same canonical keys; the reference additionally lists
    aliases of them (`_bn_modules.<i>.*`, `detection.hidden.*` = `features.layer4.*`), which are skipped."""
    sd = model.state_dict()
    with torch.no_grad():
        for k, t in sd.items():
            if k.startswith("_bn_modules.") or k.startswith("detection.hidden.") or k.startswith("features.normal."):
                continue
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(k.encode()))
            if k.endswith("num_batches_tracked"):
                t.zero_()
            elif k.endswith("running_mean"):
                t.copy_(0.1 * torch.randn(t.shape, generator=g))
            elif k.endswith("running_var"):
                t.copy_(0.5 + torch.rand(t.shape, generator=g))
            elif ("_anchor_" in k or "_proposal_" in k) and k.endswith("weight"):      # prediction layers: small, like a fresh head
                t.copy_(0.01 * torch.randn(t.shape, generator=g))
            elif t.dim() == 4:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                t.copy_(torch.randn(t.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            elif t.dim() == 2:
                t.copy_(torch.randn(t.shape, generator=g) * (1.0 / t.shape[1]) ** 0.5)
            elif k.endswith("bn3.weight"):                   # last BatchNorm of a block: small, so activations stay O(1)
                t.copy_(0.25 + 0.05 * torch.randn(t.shape, generator=g))
            elif k.endswith("weight"):                       # BatchNorm scale
                t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
            else:                                            # biases
                t.copy_(0.05 * torch.randn(t.shape, generator=g))
    return model


def make_batches(seed: int):
    g = torch.Generator().manual_seed(seed)
    images = [torch.rand(BATCH, 3, HEIGHT, WIDTH, generator=g) for _ in range(ITERS)]
    boxes, classes = [], []
    for _ in range(ITERS):
        bb = torch.zeros(BATCH, 3, 4)
        lb = torch.zeros(BATCH, 3, dtype=torch.long)
        for i, n in enumerate((3, 2)):                       # the second image has one zero-padded slot (padding_collate_fn)
            x0 = torch.rand(n, generator=g) * (WIDTH - 80)
            y0 = torch.rand(n, generator=g) * (HEIGHT - 80)
            w = 30 + torch.rand(n, generator=g) * 50
            h = 30 + torch.rand(n, generator=g) * 50
            bb[i, :n] = torch.stack((x0, y0, x0 + w, y0 + h), dim=1)
            lb[i, :n] = torch.randint(1, MODEL["num_classes"], (n,), generator=g)
        boxes.append(bb)
        classes.append(lb)
    return images, boxes, classes


def install_support_stubs(oracle_mod):
    """`support.layer.nms.nms` and `support.layer.roi_align.ROIAlign` on the C restatement (oracle/afan_oracle.c)."""

    def nms(bboxes, scores, threshold):                      # nms.cu:126-130: kept original indices, ascending
        keep = oracle_mod.nms(bboxes.detach().numpy(), scores.detach().numpy(), float(threshold), strict_gt=True)
        return torch.from_numpy(keep)

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, feat, rois, output_size, spatial_scale, sampling_ratio):
            ctx.save_for_backward(feat, rois)
            ctx.cfg = (output_size, spatial_scale, sampling_ratio)
            return torch.from_numpy(oracle_mod.roi_align(feat.detach().numpy(), rois.detach().numpy(), *ctx.cfg))

        @staticmethod
        def backward(ctx, dout):
            feat, rois = ctx.saved_tensors
            _, dfeat = oracle_mod.roi_align(feat.detach().numpy(), rois.detach().numpy(), *ctx.cfg, dout=dout.contiguous().numpy())
            return torch.from_numpy(dfeat), None, None, None, None

    class ROIAlign(nn.Module):
        def __init__(self, output_size, spatial_scale, sampling_ratio):
            super().__init__()
            self.output_size, self.spatial_scale, self.sampling_ratio = output_size, spatial_scale, sampling_ratio

        def forward(self, input, rois):
            return _Fn.apply(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    support, layer = types.ModuleType("support"), types.ModuleType("support.layer")
    nms_mod, ra_mod = types.ModuleType("support.layer.nms"), types.ModuleType("support.layer.roi_align")
    nms_mod.nms, ra_mod.ROIAlign, ra_mod.roi_align = nms, ROIAlign, _Fn.apply
    support.layer, layer.nms, layer.roi_align = layer, nms_mod, ra_mod
    for name, mod in (("support", support), ("support.layer", layer), ("support.layer.nms", nms_mod),
                      ("support.layer.roi_align", ra_mod)):
        sys.modules[name] = mod
    return nms, _Fn.apply


def _load_reference():
    try:
        from oracle import ref_shim, oracle as oracle_mod
    except ImportError:                                      # run as a script from inside oracle/
        sys.path.insert(0, os.path.dirname(HERE))
        from oracle import ref_shim, oracle as oracle_mod
    ref_shim._install_stubs()
    install_support_stubs(oracle_mod)
    det_root = os.path.join(ref_shim.REFERENCE_ROOT, "Detection")
    attack_algo = ref_shim.load("Detection", "attack_algo")
    sys.path.insert(0, det_root)
    sys.dont_write_bytecode = True
    try:
        import model as det_model                            # Detection/model.py
        import backbone.base
        import backbone.resnet101_ori as r101
        from roi.pooler import Pooler
    finally:
        sys.path.remove(det_root)
    return ref_shim, attack_algo, det_model, backbone.base, r101, Pooler


def build_reference_model(det_model, backbone_base, r101, Pooler, seed: int = 7):
    class NarrowBackbone(backbone_base.Base):                # backbone/resnet101.py:14-33 on a narrow ResNet
        def features(self):
            net = r101.ResNet(r101.Bottleneck, list(MODEL["layers"]), width_per_group=MODEL["base_width"])
            for part in (net.conv1, net.bn1, net.relu, net.maxpool, net.layer1):
                for p in part.parameters():
                    p.requires_grad = False
            return net, list(net.children())[-3], 1024, 2048

    model = det_model.Model(NarrowBackbone(pretrained=False), MODEL["num_classes"], pooler_mode=Pooler.Mode.ALIGN,
                            anchor_ratios=MODEL["anchor_ratios"], anchor_sizes=MODEL["anchor_sizes"],
                            rpn_pre_nms_top_n=MODEL["pre_nms"], rpn_post_nms_top_n=MODEL["post_nms"],
                            anchor_smooth_l1_loss_beta=MODEL["beta"], proposal_smooth_l1_loss_beta=MODEL["beta"])
    return procedural_init(model, seed)


def reference_iteration(model, attack_algo, image_batch, bboxes_batch, labels_batch, c, optimizer):
    """train_aug_final.py:78-163, one pass of the loop body (nn.DataParallel on one device is the identity)."""
    f1, f2, f3, f4 = (int(ch) for ch in c["mix_layer"])                                                  # :75-76
    y = {"bb": bboxes_batch, "lb": labels_batch}
    inputs_all_se = {"x": image_batch, "adv": None, "out_idx": c["se"], "flag": "head"}                  # :83
    inputs_all_sd = {"x": image_batch, "adv": None, "out_idx": "roi_head", "flag": "clean"}              # :84
    feature_map_se = model.train().forward(inputs_all_se, bboxes_batch, labels_batch).detach()           # :86-87
    rpn_roi_output_dict = model.train().forward(inputs_all_sd, bboxes_batch, labels_batch)               # :89
    clean_feature_map_sd = rpn_roi_output_dict["roi_output_dict"]["roi_feature_map"].detach()            # :90
    feature_adv_se = attack_algo.PGD(feature_map_se, image_batch, y=y, model=model, steps=1, eps=(2.0 / 255),
                                     gamma=(c["gamma_se"] / 255), idx=c["se"], randinit=c["randinit"], clip=c["clip"])   # :92-100
    adv_rpn_roi_output_dict = attack_algo.rpn_roi_PGD(layer="roi", rpn_roi_output_dict=rpn_roi_output_dict, y=y, model=model,
                                                      steps=1, eps=(2.0 / 255), gamma=(c["gamma_sd"] / 255),
                                                      randinit=c["randinit"], clip=c["clip"],
                                                      only_roi_loss=c["only_roi_sd"])                    # :102-112
    adv_feature_map_sd = adv_rpn_roi_output_dict["roi_output_dict"]["roi_feature_map"].detach()         # :114
    if c["mix_sd"]:
        adv_feature_map_sd = attack_algo.mix_feature(clean_feature_map_sd, adv_feature_map_sd)           # :116-117
    if c["noise_sd"] != 0:
        adv_feature_map_sd += (2.0 * torch.rand(adv_feature_map_sd.shape).cuda() - 1.0) * c["gamma_sd"] * c["noise_sd"]   # :118-119
    adv_rpn_roi_output_dict["roi_output_dict"]["roi_feature_map"] = adv_feature_map_sd                   # :120
    adv_list_se = attack_algo.get_sample_points(feature_map_se, feature_adv_se, 5)                       # :122
    for i, f in ((1, f1), (2, f2), (3, f3), (4, f4)):                                                    # :124-131
        if f:
            adv_list_se[i] = attack_algo.mix_feature(feature_map_se, adv_list_se[i])
    dicts = [{"x": image_batch, "adv": None, "out_idx": 0, "flag": "clean"}]                             # :133
    dicts += [{"x": image_batch, "adv": adv_list_se[i], "out_idx": c["se"], "flag": "tail"} for i in (1, 2, 3, 4)]   # :134-137
    dicts += [{"adv": adv_rpn_roi_output_dict, "out_idx": "roi_tail", "flag": "clean"}]                  # :138
    losses = [attack_algo.compute_loss(*model.train().forward(d, bboxes_batch, labels_batch)) for d in dicts]       # :140-158
    loss = (sum(losses[:5]) / 3.0) * (1 - c["w"]) + (losses[5] / 3.0) * c["w"]                           # :161
    optimizer.zero_grad()                                                                                # :166
    loss.backward()
    optimizer.step()
    return [float(l) for l in losses] + [float(loss)]


def generate():
    ref_shim, attack_algo, det_model, backbone_base, r101, Pooler = _load_reference()
    out = {}
    for name, c in CASES.items():
        with ref_shim.cpu_cuda_identity():
            model = build_reference_model(det_model, backbone_base, r101, Pooler)
            optimizer = torch.optim.SGD(model.parameters(), lr=LR, momentum=MOMENTUM, weight_decay=WD)    # :46-47
            images, boxes, classes = make_batches(seed=33)
            losses = []
            for it in range(ITERS):
                torch.manual_seed(200 + it)
                losses.append(reference_iteration(model, attack_algo, images[it], boxes[it], classes[it], c, optimizer))
                print(name, it, losses[-1], flush=True)
        out[f"{name}/losses"] = np.array(losses, dtype=np.float64)
        sd = model.state_dict()
        keys = [k for k in sd.keys() if not (k.startswith("_bn_modules.") or k.startswith("detection.hidden."))]
        out[f"{name}/norms"] = np.array([float(sd[k].double().norm()) for k in keys], dtype=np.float64)
        for k in FULL:
            out[f"{name}/final/{k}"] = sd[k].detach().numpy().copy()
    out["keys"] = np.array(keys)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    generate()
