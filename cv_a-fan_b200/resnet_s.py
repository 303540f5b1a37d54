"""Splittable CIFAR ResNet for A-FAN, built on the fused dual-BN kernels.

Interface parity with the reference's Classification/resnet_s.py:79-124: the network is ONE
`nn.Sequential` (`sequential_model`) so `model(x, end_point=k, start_point=j)` runs layers [j, k)
(head = [0, k), tail = [k, L)); index layout, parameter / buffer names and the learnable `w` vector
match, so reference checkpoints load with `load_state_dict`.  Differences are internal: every
BatchNorm (+ the ReLU / residual add that follows it, resnet_s.py:70-76) is ONE fused kernel pair, and
forward takes two extra keyword arguments threaded to the BN layers:
    groups  statistic groups along the batch ([adv; clean] -> 2), see dual_bn.DualBatchNorm2d
    replay  how many times the running statistics are advanced (head cache: 2, main_perturb.py:173,196)
"""
import os
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import conv as conv_mod
from . import ops
from .conv import Conv3x3
from .dual_bn import DualBatchNorm2d

# AFAN_FUSE_BN1 (default on): in passes that need no weight gradients (the PGD ascent: attack_algo.py:52 `only_inputs=True`) an
# identity BasicBlock runs conv1 -> [bn1 + relu folded into conv2's operand staging] -> conv2: conv1's tcgen05 kernel reduces
# the BatchNorm statistics of its own output in its epilogue, conv2 normalises while it loads, and relu(bn1(conv1(x))) is
# never written to memory -- one BatchNorm launch and 8 B/element of traffic less per block and pass (80 launches per
# step at config 2; same-box A/B 11.18 -> 10.89 ms).  Multi-GPU: with the fused NVLink exchange (bn_exchange="p2p") the
# consumer convolution exchanges the folded sums with its peers in its prologue; with the NCCL split form the block keeps
# its four launches.
FUSE_BN1 = os.environ.get("AFAN_FUSE_BN1", "1") == "1"
# AFAN_TAIL_KERNELS=0: option-A shortcut and classifier weight gradient back on the library's launches (A/B)
SHORTCUT_KERNEL = LINEAR_KERNEL = os.environ.get("AFAN_TAIL_KERNELS", "1") == "1"
fused_forward_calls = 0       # host-side count of folded block forwards (eager calls + graph captures): evidence for the parity checks


class _FusedConvBnReluConvFn(torch.autograd.Function):
    """(conv2(relu(bn1(conv1(x)))), x) for an identity BasicBlock with frozen parameters (resnet_s.py:70-73); the second
    output aliases x for the shortcut so that its gradient joins conv1's input gradient in the dgrad epilogue."""

    @staticmethod
    def forward(ctx, x, block, groups, replay):
        x = x.contiguous()
        bn = block.bn1
        wf1, wd1 = block.conv1.packed()
        wf2, wd2 = block.conv2.packed()
        n, c = x.shape[0], x.shape[1]
        ws = block._fuse_ws.get((n, c, x.device))
        if ws is None:
            ws = block._fuse_ws[(n, c, x.device)] = ops.conv3x3_umma_bn_workspace(n, c, x.device)
        c1, _, _, _ = ops.conv3x3_umma_bn(x, wf1, stats_out=ws, groups=groups)
        c2, sm, si, tab = ops.conv3x3_umma_bn(c1, wf2, stats_in=ws, bn=(bn.weight, bn.bias, bn.running_mean, bn.running_var),
                                              groups=groups, eps=bn.eps, momentum=bn.momentum, replay=replay, mailbox=bn.mailbox)
        ctx.save_for_backward(c1, tab, sm, si, bn.weight)
        ctx.wd1, ctx.wd2, ctx.groups, ctx.mailbox = wd1, wd2, groups, bn.mailbox
        return c2, x.view_as(x)

    @staticmethod
    def backward(ctx, d_c2, dtap):
        c1, tab, sm, si, w = ctx.saved_tensors
        d_h = ops.conv3x3(d_c2.contiguous(), ctx.wd2, math="umma")
        d_c1, _, _ = ops.bn_bwd_xmask(d_h, c1, tab, w, sm, si, groups=ctx.groups, mailbox=ctx.mailbox)
        dx = ops.conv3x3(d_c1, ctx.wd1, math="umma", addend=dtap.contiguous() if dtap is not None else None)
        return dx, None, None, None

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)
CIFAR_STD = (0.2470, 0.2435, 0.2616)


class NormalizeByChannelMeanStd(nn.Module):
    """Input normalisation layer 0 of the Sequential (advertorch's module in the reference, resnet_s.py:87)."""

    def __init__(self, mean: Sequence[float], std: Sequence[float]):
        super().__init__()
        self.register_buffer("mean", torch.tensor(mean, dtype=torch.float32))
        self.register_buffer("std", torch.tensor(std, dtype=torch.float32))

    def forward(self, x):
        return (x - self.mean[None, :, None, None]) / self.std[None, :, None, None]


class BasicBlock(nn.Module):
    """conv-bn-relu-conv-bn-(+shortcut)-relu with option-A shortcut (resnet_s.py:45-77).  bn1+relu and
    bn2+add+relu are each one fused dual-BN call."""
    expansion = 1

    def __init__(self, in_planes: int, planes: int, stride: int = 1):
        super().__init__()
        self.conv1 = Conv3x3(in_planes, planes, stride)
        self.bn1 = DualBatchNorm2d(planes)
        self.conv2 = Conv3x3(planes, planes, 1)
        self.bn2 = DualBatchNorm2d(planes)
        self.shortcut = nn.Sequential()            # keeps the (parameter-free) child name of the reference
        self.downsample = stride != 1 or in_planes != planes
        self.pad = planes // 4
        self._fuse_ws = {}

    def _fusable(self, x, groups: int) -> bool:
        if not (FUSE_BN1 and self.training and not self.downsample and torch.is_grad_enabled() and x.requires_grad):
            return False
        if conv_mod.MODE != "tc3" or groups > 2 or x.shape[0] % groups:
            return False
        n, c, h = x.shape[0], x.shape[1], x.shape[2]
        if not (x.is_cuda and x.dtype == torch.float32 and ops.conv3x3_umma_supported(n, c, h)):
            return False
        if h == 8 and (n // groups) % 2:
            return False
        if self.bn1.mailbox is not None:          # multi-GPU: the exchange runs inside the consumer convolution; every CTA
            if n > torch.cuda.get_device_properties(x.device).multi_processor_count:      # spins on its peers -> co-resident grid
                return False
        elif self.bn1.process_group is not None:  # NCCL split form of the statistics exchange: keep the four launches
            return False
        return not any(p.requires_grad for m in (self.conv1, self.bn1, self.conv2) for p in m.parameters())

    def _shortcut(self, x):
        if not self.downsample:
            return x
        if x.is_cuda and x.dtype == torch.float32 and SHORTCUT_KERNEL:
            return ops.shortcut_a(x, self.pad)                       # one launch each way instead of slice + fill + pad
        return F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, self.pad, self.pad), "constant", 0.0)

    def forward(self, x, groups: int = 1, replay: int = 1):
        if self._fusable(x, groups):
            global fused_forward_calls
            fused_forward_calls += 1
            c2, sc = _FusedConvBnReluConvFn.apply(x, self, groups, replay)
            self.bn1._pending_batches += groups * replay
            return self.bn2(c2, residual=sc, relu=True, groups=groups, replay=replay)
        if self.downsample:
            c1, sc = self.conv1(x), self._shortcut(x)
        else:
            c1, sc = self.conv1.forward_with_tap(x)      # identity shortcut: its gradient joins conv1's dgrad in-kernel
        h = self.bn1(c1, relu=True, groups=groups, replay=replay)
        return self.bn2(self.conv2(h), residual=sc, relu=True, groups=groups, replay=replay)


class ResNet(nn.Module):
    def __init__(self, block=BasicBlock, num_blocks=(9, 9, 9), num_classes: int = 10, init_weight: float = 1.0):
        super().__init__()
        layers = [NormalizeByChannelMeanStd(CIFAR_MEAN, CIFAR_STD),
                  nn.Conv2d(3, 16, 3, 1, 1, bias=False), DualBatchNorm2d(16), nn.ReLU()]
        in_planes = 16
        for planes, stride, depth in zip((16, 32, 64), (1, 2, 2), num_blocks):
            for i in range(depth):
                layers.append(block(in_planes, planes, stride if i == 0 else 1))
                in_planes = planes * block.expansion
        layers += [nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(), nn.Linear(64, num_classes)]
        self.sequential_model = nn.Sequential(*layers)
        self.all_layers = 9
        self.w = nn.Parameter(torch.full((self.all_layers,), float(init_weight)))
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight)

    @property
    def layer_number(self) -> int:
        return len(self.sequential_model)

    def forward(self, x, end_point: Optional[int] = None, start_point: int = 0, groups: int = 1, replay: int = 1):
        layers = self.sequential_model
        end_point = len(layers) if end_point is None else min(end_point, len(layers))
        i = start_point
        while i < end_point:
            m = layers[i]
            if isinstance(m, DualBatchNorm2d):
                fuse = i + 1 < end_point and isinstance(layers[i + 1], nn.ReLU)     # stem bn + relu in one kernel pair
                x = m(x, relu=fuse, groups=groups, replay=replay)
                i += 2 if fuse else 1
                continue
            if isinstance(m, BasicBlock):
                x = m(x, groups=groups, replay=replay)
            elif isinstance(m, nn.Linear) and LINEAR_KERNEL and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2:
                x = ops.linear(x, m.weight, m.bias)                  # weight / bias gradient in one deterministic launch
            else:
                x = m(x)
            i += 1
        return x

    def bn_layers(self):
        return [m for m in self.modules() if isinstance(m, DualBatchNorm2d)]


def resnet20(num_classes: int = 10, **kw):
    return ResNet(BasicBlock, (3, 3, 3), num_classes, **kw)


def resnet56(init_weight_eta: float = 1.0, num_classes: int = 10):
    """Factory named like the reference's (resnet_s.py:123-124); num_classes=100 for CIFAR-100."""
    return ResNet(BasicBlock, (9, 9, 9), num_classes, init_weight=init_weight_eta)
