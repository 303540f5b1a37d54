"""-m gpu: the BASELINE.json configs that are parity-test cases rather than bench lines (configs 3-5): the drop-in PGD
of each flavour driven by a real (library, cuDNN) tail network at the config's feature shape, checked against the
reference's update arithmetic applied to the SAME gradients (recomputed with plain torch ops on the GPU).
Gradients come from cuDNN in both runs; a handful of near-zero gradients may differ in sign between two backward
passes (non-deterministic reductions), so bit-equality is required on all but <= 0.01 % of the elements."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.util import PKG, dev

pytestmark = pytest.mark.gpu


def reference_update(x_adv, g, x, gamma, eps, clip):
    """Classification/attack_algo.py:53-56 as the single expression proven bit-identical to it."""
    t = x_adv + gamma * torch.sign(g)
    if clip:
        lo, hi = x - eps, x + eps
        t = torch.where(t < lo, lo, t)
        t = torch.where(t > hi, hi, t)
    return t


def mismatch(a, b):
    return float((a != b).float().mean())


@pytest.fixture(autouse=True)
def strict():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_config3_efficientnet_b0_bf16_feature_pgd1():
    """EfficientNet-B0, 224x224, PGD-1 on the features[3] output (N x 40 x 28 x 28), bf16 storage."""
    tv = pytest.importorskip("torchvision")
    torch.manual_seed(0)
    net = tv.models.efficientnet_b0(num_classes=10).to(dev()).bfloat16().eval()      # eval: frozen BN like a fine-tune tail

    def model(x, end_point=None, start_point=0):          # the reference's split contract on torchvision's Sequential
        if start_point == 0:
            return net.features[:end_point](x)
        return net.classifier(torch.flatten(net.avgpool(net.features[start_point:](x)), 1))

    g = torch.Generator().manual_seed(1)
    imgs = torch.rand(8, 3, 224, 224, generator=g).to(dev()).bfloat16()
    y = torch.randint(0, 10, (8,), generator=g).to(dev())
    with torch.no_grad():
        feat = model(imgs, end_point=4).contiguous()
    assert feat.shape == (8, 40, 28, 28) and feat.dtype == torch.bfloat16
    u = torch.rand(feat.shape, generator=g)
    gamma, eps = 1.0 / 255, 2.0 / 255
    ce = nn.CrossEntropyLoss()
    x_adv = PKG.attack_algo.PGD(feat, ce, y=y, model=model, steps=1, gamma=gamma, start_idx=4, layer_number=9, eps=eps,
                                randinit=True, clip=True, noise=u)
    assert x_adv.dtype == torch.bfloat16 and x_adv.is_leaf and x_adv.requires_grad
    # expected: fp32 arithmetic on the widened values, rounded to bf16 (the bf16 twin's contract)
    x0 = (feat.float() + (2.0 * u.to(dev()) - 1.0) * eps).bfloat16()
    xa = x0.clone().requires_grad_(True)
    grad = torch.autograd.grad(ce(model(xa, start_point=4), y), xa)[0]
    exp = reference_update(x0.float(), grad.float(), feat.float(), gamma, eps, True).bfloat16()
    assert mismatch(x_adv.detach(), exp) <= 1e-4
    assert float((x_adv.detach().float() - feat.float()).abs().max()) <= eps * 1.02 + 2 ** -8 * float(feat.float().abs().max())


class _SegStandIn(nn.Module):
    """Dict-API tail like Segmentation/network/utils.py:14-46 ('flag': 'tail'): perturbed layer-k feature ->
    classifier -> bilinear upsample to the image size."""

    def __init__(self, c_in, n_cls):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(c_in, 64, 3, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU(),
                                  nn.Conv2d(64, n_cls, 1))

    def forward(self, inputs):
        assert inputs["flag"] == "tail"
        out = self.body(inputs["adv"])
        return F.interpolate(out, size=inputs["x"].shape[-2:], mode="bilinear", align_corners=False)


def test_config5_deeplab_shaped_seg_flavour():
    """DeepLabv3+ shape: ASPP-input feature 2 x 2048 x 33 x 33 (513^2 crops, os16), Seg flavour, 2 steps + clip."""
    torch.manual_seed(0)
    model = _SegStandIn(2048, 19).to(dev()).train()
    g = torch.Generator().manual_seed(2)
    images = torch.rand(2, 3, 513, 513, generator=g).to(dev())
    labels = torch.randint(0, 19, (2, 513, 513), generator=g).to(dev())
    x = torch.relu(torch.randn(2, 2048, 33, 33, generator=g)).to(dev())
    crit = nn.CrossEntropyLoss(ignore_index=255)
    gamma, eps = 0.01 / 255 * 100, 2.0 / 255
    out = PKG.segmentation.PGD(x, images, None, crit, y=labels, model=model, steps=2, eps=eps, gamma=gamma, idx=4,
                               randinit=False, clip=True)
    xa = x.clone()
    for _ in range(2):
        xr = xa.clone().requires_grad_(True)
        gr = torch.autograd.grad(crit(model({"x": images, "adv": xr, "out_idx": 4, "flag": "tail", "low_level_feat": None}), labels), xr)[0]
        xa = reference_update(xa, gr, x, gamma, eps, True)
    assert mismatch(out.detach(), xa) <= 1e-4
    mixed = PKG.segmentation.mix_feature(x, out.detach())
    ref = ((x - x.mean(1, keepdim=True)) / (x.var(1, keepdim=True) + 1e-5).sqrt()
           * (out.detach().var(1, keepdim=True) + 1e-5).sqrt() + out.detach().mean(1, keepdim=True))
    torch.testing.assert_close(mixed, ref, rtol=2e-5, atol=2e-6)


class _DetStandIn(nn.Module):
    """Detection contract: model.train().forward(inputs, bb, lb) -> 4 loss tensors (Detection/model.py:58-75)."""

    def __init__(self, c_in):
        super().__init__()
        self.rpn = nn.Conv2d(c_in, 32, 3, padding=1)
        self.head = nn.Linear(32, 4)

    def forward(self, inputs, bb, lb):
        h = F.relu(self.rpn(inputs["adv"]))
        pooled = self.head(h.mean(dim=(2, 3)))
        return pooled[:, 0].abs(), (h ** 2).mean(dim=(1, 2, 3)), pooled[:, 1:3].pow(2).sum(1), F.smooth_l1_loss(pooled[:, 3], bb, reduction="none")


def test_config4_faster_rcnn_shaped_det_flavour():
    """Faster R-CNN shape: backbone layer3 feature B x 1024 x 38 x 63 (600x1000 inputs), Det flavour PGD-1, no clip."""
    torch.manual_seed(0)
    model = _DetStandIn(1024).to(dev())
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn(2, 1024, 38, 63, generator=g)).to(dev())
    bb = torch.randn(2, generator=g).to(dev())
    gamma, eps = 0.5 / 255, 2.0 / 255
    out = PKG.detection.PGD(x, None, y={"bb": bb, "lb": None}, model=model, steps=1, eps=eps, gamma=gamma, idx=3,
                            randinit=False, clip=False)
    xr = x.clone().requires_grad_(True)
    l = model.train().forward({"x": None, "adv": xr, "out_idx": 3, "flag": "tail"}, bb, None)
    gr = torch.autograd.grad(PKG.detection.compute_loss(*l), xr)[0]
    assert mismatch(out.detach(), reference_update(x, gr, x, gamma, eps, False)) <= 1e-4
