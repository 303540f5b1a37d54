"""Plain-PyTorch (CPU, un-fused ATen ops) restatement of the reference's A-FAN training path.

TEST INFRASTRUCTURE ONLY -- this is the end-to-end ORACLE and the "port" CPU baseline:
imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
The product package (cv_a-fan_b200/) never imports it.

What is restated (reference file:line, relative to the reference root):
  * CifarResNetRef      <- Classification/resnet_s.py:34-124 (option-A BasicBlock net laid out
                           as ONE nn.Sequential so that model(x, end_point, start_point) runs a
                           slice; parameter names are kept so reference state_dicts load)
  * pgd_reference       <- Classification/attack_algo.py:38-58 (+ tensor_clamp :9-19)
  * afan_train_iteration<- Classification/main_perturb.py:173-201 (head fwd, PGD, delta norms,
                           adv tail fwd, full clean fwd, mean of two CE, SGD step)
Parity pin: tests/golden/cls_train_*.npz hold losses / weights produced by executing the
unmodified reference `main_perturb.train` under oracle/ref_shim.py; tests/test_oracle_golden.py
replays them through this file (bitwise on CPU, same op order).
"""
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)
CIFAR_STD = (0.2470, 0.2435, 0.2616)


class ChannelNormalize(nn.Module):
    """advertorch.utils.NormalizeByChannelMeanStd as used at resnet_s.py:87."""

    def __init__(self, mean: Sequence[float], std: Sequence[float]):
        super().__init__()
        self.register_buffer("mean", torch.tensor(mean, dtype=torch.float32))
        self.register_buffer("std", torch.tensor(std, dtype=torch.float32))

    def forward(self, x):
        return (x - self.mean[None, :, None, None]) / self.std[None, :, None, None]


class OptionAShortcut(nn.Module):
    """resnet_s.py:63-64: spatial stride-2 subsample, zero-pad planes//4 channels each side."""

    def __init__(self, pad: int):
        super().__init__()
        self.pad = pad

    def forward(self, x):
        return F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, self.pad, self.pad), "constant", 0)


class ResidualUnitRef(nn.Module):
    """resnet_s.py:45-77 (option A only; that is the one the reference instantiates)."""

    def __init__(self, c_in: int, c_out: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(c_in, c_out, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(c_out)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(c_out)
        self.shortcut = OptionAShortcut(c_out // 4) if (stride != 1 or c_in != c_out) else nn.Sequential()

    def forward(self, x):
        h = F.relu(self.bn1(self.conv1(x)))
        h = self.bn2(self.conv2(h))
        h = h + self.shortcut(x)
        return F.relu(h)


class CifarResNetRef(nn.Module):
    """resnet_s.py:79-121.  depth spec e.g. (9,9,9) = ResNet-56, (3,3,3) = ResNet-20."""

    def __init__(self, num_blocks=(9, 9, 9), num_classes: int = 10, init_weight: float = 1.0):
        super().__init__()
        layers: List[nn.Module] = [ChannelNormalize(CIFAR_MEAN, CIFAR_STD),
                                   nn.Conv2d(3, 16, 3, 1, 1, bias=False), nn.BatchNorm2d(16), nn.ReLU()]
        width_in = 16
        for stage, (width, first_stride) in enumerate(((16, 1), (32, 2), (64, 2))):
            for b in range(num_blocks[stage]):
                layers.append(ResidualUnitRef(width_in, width, first_stride if b == 0 else 1))
                width_in = width
        layers += [nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(), nn.Linear(64, num_classes)]
        self.sequential_model = nn.Sequential(*layers)
        self.w = nn.Parameter(torch.full((9,), float(init_weight)))      # resnet_s.py:113-114
        for m in self.modules():                                           # resnet_s.py:36-40,116
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                nn.init.kaiming_normal_(m.weight)

    @property
    def layer_number(self) -> int:
        return len(self.sequential_model)

    def forward(self, x, end_point: Optional[int] = None, start_point: int = 0):
        end_point = len(self.sequential_model) if end_point is None else end_point
        return self.sequential_model[start_point:end_point](x)


def pgd_reference(x, loss_fn, y, model, steps, gamma, start_idx, layer_number, eps, randinit, clip,
                  noise=None):
    """Classification/attack_algo.py:38-58 in un-fused ATen ops.  `noise` = the torch.rand draw
    (None -> drawn here from the CPU generator exactly where the reference draws it)."""
    x_adv = x.clone()
    if randinit:
        u = torch.rand(x_adv.shape) if noise is None else noise
        x_adv += (2.0 * u.to(x_adv.device) - 1.0) * eps
    x_adv.requires_grad_(True)
    for _ in range(steps):
        out = model(x_adv, end_point=layer_number, start_point=start_idx)
        loss = loss_fn(out, y)
        g = torch.autograd.grad(loss, x_adv, only_inputs=True)[0]
        x_adv.data.add_(gamma * torch.sign(g.data))
        if clip:
            lo, hi = x - eps, x + eps
            t = x_adv.data
            t.copy_(torch.where(t < lo, lo, t))
            t.copy_(torch.where(t > hi, hi, t))
    return x_adv


def afan_train_iteration(model, optimizer, criterion, images, target, *, steps, gamma, eps, perturb_idx,
                         randinit=False, clip=False, noise=None):
    """Body of Classification/main_perturb.py:173-201.  gamma/eps are in 1/255 units like the
    reference flags.  Returns (loss, output_clean, l2[N], linf[N], feature_adv)."""
    layer_number = len(model.sequential_model)
    feature = model(images, end_point=perturb_idx, start_point=0).detach()
    feature_adv = pgd_reference(feature, criterion, target, model, steps, gamma / 255, perturb_idx,
                                layer_number, eps / 255, randinit, clip, noise=noise)
    delta = (feature_adv - feature).detach().reshape(images.shape[0], -1)
    l2 = torch.norm(delta, p=2, dim=1)
    linf = torch.norm(delta, p=float("inf"), dim=1)
    out_adv = model(feature_adv, end_point=layer_number, start_point=perturb_idx)
    out_clean = model(images, end_point=layer_number, start_point=0)
    loss = (criterion(out_adv, target) + criterion(out_clean, target)) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach(), out_clean.detach(), l2, linf, feature_adv.detach()


def make_sgd(model, lr=0.1, momentum=0.9, weight_decay=5e-4):
    """main_perturb.py:72-74."""
    return torch.optim.SGD(model.parameters(), lr, momentum=momentum, weight_decay=weight_decay)


def mix_feature_reference(clean, adv):
    """Segmentation/attack_algo.py:121-130 in torch ops (channel-dim mean / unbiased var)."""
    e = 1e-5
    m_c, m_a = clean.mean(1, keepdim=True), adv.mean(1, keepdim=True)
    s_c = (clean.var(1, keepdim=True) + e).sqrt()
    s_a = (adv.var(1, keepdim=True) + e).sqrt()
    return (clean - m_c) / s_c * s_a + m_a
