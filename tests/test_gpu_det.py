"""Rows a6 / f3 (Detection flavour): cv_a-fan_b200.faster_rcnn + trainer_det.DetAfanTrainer against
tests/golden/det_step.npz, produced by executing the reference's training-iteration body
(Detection/train_aug_final.py:78-163, restated around the unmodified reference Model / RPN / attack_algo in
oracle/det_ref_step.py; `support._C` supplied by the C restatement of nms.cu / ROIAlign_cuda.cu) on the CPU.

The GPU run replays the same inputs, the same key-keyed initial weights and the SAME CPU random stream
(torch.manual_seed(200 + it): candidate sampling via Sampler('reference'), random starts and noise via rng='reference').
Tolerances: the reference ran mkldnn fp32 on the CPU, this runs cuDNN fp32 + the sm_100a kernels; every discrete decision
(labels, NMS, sampling) must come out the same for the losses to agree at 2e-3 relative."""
import importlib

import numpy as np
import pytest
import torch

from oracle import det_ref_step as ref
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PKG = importlib.import_module("cv_a-fan_b200")
G = np.load(ref.GOLDEN, allow_pickle=False)


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _model(dev, sampler="reference", fuse=True):
    m = ref.MODEL
    model = PKG.faster_rcnn.FasterRCNN(m["num_classes"], anchor_ratios=m["anchor_ratios"], anchor_sizes=m["anchor_sizes"],
                                       rpn_pre_nms_top_n=m["pre_nms"], rpn_post_nms_top_n=m["post_nms"], layers=m["layers"],
                                       base_width=m["base_width"], sampler=sampler, fuse_frozen_bn=fuse)
    return ref.procedural_init(model, 7).to(dev)


def _run(name, head_cache, fuse=True):
    c = ref.CASES[name]
    dev = torch.device("cuda:0")
    model = _model(dev, fuse=fuse)
    tr = PKG.trainer_det.DetAfanTrainer(model, pertub_idx_se=c["se"], gamma_se=c["gamma_se"], gamma_sd=c["gamma_sd"],
                                        randinit=c["randinit"], clip=c["clip"], mix_layer=c["mix_layer"], noise_sd=c["noise_sd"],
                                        only_roi_sd=c["only_roi_sd"], mix_sd=c["mix_sd"], sd_adv_loss_weight=c["w"], lr=ref.LR,
                                        momentum=ref.MOMENTUM, weight_decay=ref.WD, head_cache=head_cache, rng="reference")
    images, boxes, classes = ref.make_batches(seed=33)
    losses = []
    for it in range(ref.ITERS):
        torch.manual_seed(200 + it)
        out = tr.step(images[it].to(dev), boxes[it].to(dev), classes[it].to(dev))
        losses.append(out["losses"].cpu().tolist() + [float(out["loss"])])
    return np.array(losses), {k: v.detach().float().cpu() for k, v in model.state_dict().items()}


@pytest.mark.parametrize("head_cache,fuse", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("name", ["A", "B"])
def test_detection_iteration_matches_executed_reference(name, head_cache, fuse):
    losses, sd = _run(name, head_cache, fuse)
    np.testing.assert_allclose(losses, G[f"{name}/losses"], rtol=2e-3, atol=1e-5)
    keys = [str(k) for k in G["keys"]]
    for k, gold in zip(keys, G[f"{name}/norms"]):
        if k.startswith("features.normal."):
            continue
        assert abs(float(sd[k].double().norm()) - gold) <= 2e-4 * max(gold, 1.0), k
    for k in ref.FULL:
        torch.testing.assert_close(sd[k], torch.from_numpy(G[f"{name}/final/{k}"]), rtol=0, atol=2e-4)
    frozen = [k for k in keys if ".bn" in k or k.startswith("features.conv1") or k.startswith("features.layer1")]
    init = ref.procedural_init(_model("cpu"), 7).state_dict()
    for k in frozen:                                          # frozen BatchNorm (statistics AND affine), conv1, layer1
        assert torch.equal(sd[k], init[k].float()), k


def test_head_cache_is_exact():
    """One backbone sweep / one RPN prediction / one NMS for the three forwards that repeat them (trainer_det.py): same
    losses and weights as the literal schedule, to cuDNN run-to-run determinism."""
    a, sa = _run("A", True)
    b, sb = _run("A", False)
    np.testing.assert_allclose(a, b, rtol=1e-5)
    for k in sa:
        torch.testing.assert_close(sa[k], sb[k], rtol=1e-4, atol=1e-6)


def test_proposals_and_roi_pooling_match_the_c_oracle():
    """The two native kernels inside the model, at the model's own operating point: proposals (decode + clip + rank + NMS on
    the device) against the C restatement of nms.cu, ROIAlign forward / backward against that of ROIAlign_cuda.cu."""
    dev = torch.device("cuda:0")
    model = _model(dev).train()
    images, boxes, classes = ref.make_batches(seed=33)
    x = images[0].to(dev)
    feats = model.features({"x": x, "flag": "clean", "out_idx": 0})
    r = model.rpn_outputs(x, feats)
    obj, trf, anchors = r["objectnesses"].detach(), r["transformers"].detach(), r["anchors"]
    dec = PKG.faster_rcnn.clip_boxes(PKG.faster_rcnn.apply_deltas(anchors.unsqueeze(0), trf), x.shape[3], x.shape[2])
    for i in range(x.shape[0]):
        score, order = torch.sort(obj[i, :, 1], descending=True, stable=True)
        ranked = dec[i][order][:ref.MODEL["pre_nms"]].cpu().numpy()
        keep = orc.nms(ranked, score[:ref.MODEL["pre_nms"]].cpu().numpy(), 0.7, strict_gt=True)[:ref.MODEL["post_nms"]]
        got = r["proposals"][i].cpu().numpy()
        assert np.array_equal(got[:len(keep)], ranked[keep])
        assert not got[len(keep):].any()
    rois = torch.cat((torch.zeros(8, 1, device=dev), r["proposals"][0, :8]), dim=1)
    f = feats.detach().clone().requires_grad_(True)
    out = PKG.detection.roi_align(f, rois, (14, 14), 1 / 16, 0)
    dout = torch.randn_like(out)
    out.backward(dout)
    want, dwant = orc.roi_align(feats.detach().cpu().numpy(), rois.cpu().numpy(), (14, 14), 1 / 16, 0, dout=dout.cpu().numpy())
    np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(f.grad.cpu().numpy(), dwant, rtol=1e-4, atol=1e-4)


def test_device_sampler_runs_and_trains():
    """Sampler('device') + rng='philox': no CPU random stream at all; the loss must be finite and the weights must move."""
    dev = torch.device("cuda:0")
    model = _model(dev, sampler="device")
    before = model.rpn._anchor_objectness.weight.detach().clone()
    tr = PKG.trainer_det.DetAfanTrainer(model, pertub_idx_se=2, randinit=True, clip=True, mix_layer="0110", mix_sd=True, noise_sd=0.5,
                                        lr=ref.LR, rng="philox", seed=5)
    images, boxes, classes = ref.make_batches(seed=33)
    for it in range(2):
        out = tr.step(images[it].to(dev), boxes[it].to(dev), classes[it].to(dev))
        assert torch.isfinite(out["losses"]).all()
    assert not torch.equal(before, model.rpn._anchor_objectness.weight.detach())
