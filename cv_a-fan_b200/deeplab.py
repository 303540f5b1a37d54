"""Splittable DeepLabv3+ for the Segmentation flavour of A-FAN (row f3).

Interface parity with the reference's Segmentation/network (`_SimpleSegmentationModel.forward`, network/utils.py:14-46;
backbone dict protocol, network/backbone/resnet.py:198-304; head `return_type`s, network/_deeplab.py:46-80): the model is
called with ONE dict

    {'x': images, 'adv': feature | dict | None, 'out_idx': 1..4 | 'aspp_head' | 'concat_head' | 'aspp_tail' |
     'concat_tail' | 0, 'flag': 'head' | 'tail' | 'clean', 'low_level_feat': tensor}

and parameter / buffer names equal the reference's (`backbone.normal.mean`, `backbone.layer3.0.conv1.weight`,
`classifier.aspp.convs.1.0.weight`, ...), so its checkpoints load with `load_state_dict`.  The backbone is torchvision's
ResNet (library model code: the reference's own class is that file's 2019 copy) with `replace_stride_with_dilation`
chosen from the output stride exactly like network/modeling.py:8-13; convolutions and BatchNorm stay library kernels
(cuDNN) -- what is hand-written on this path is the perturbation loop around the model.

Extra, for the head cache of `trainer_seg.SegAfanTrainer`: `stages(x)` returns every stage output of ONE backbone sweep,
`run_tail(feature, idx)` continues from stage `idx`, and `decode(...)` exposes the head's pieces.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

from ._lib import AfanError
from .resnet_s import NormalizeByChannelMeanStd

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class SplitResNet(nn.Module):
    """ResNet-50/101 trunk with the reference's child names (normal, conv1, bn1, relu, maxpool, layer1..4)."""

    def __init__(self, name: str = "resnet101", output_stride: int = 16):
        super().__init__()
        dil = [False, True, True] if output_stride == 8 else [False, False, True]      # network/modeling.py:8-13
        tv = getattr(torchvision.models, name)(weights=None, replace_stride_with_dilation=dil)
        self.normal = NormalizeByChannelMeanStd(IMAGENET_MEAN, IMAGENET_STD)
        for child in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4"):
            setattr(self, child, getattr(tv, child))

    def stem(self, x):
        return self.layer1(self.maxpool(self.relu(self.bn1(self.conv1(self.normal(x))))))

    def run_from(self, x, idx: int):
        """Continue after stage `idx` (1 = output of layer1 ... 4 = output of layer4) to the end of layer4."""
        for k in range(idx + 1, 5):
            x = getattr(self, f"layer{k}")(x)
        return x

    def stages(self, x, upto: int = 4):
        """{1: layer1 out (= low_level), 2: ..., upto: ...} from one sweep."""
        out = {1: self.stem(x)}
        for k in range(2, upto + 1):
            out[k] = getattr(self, f"layer{k}")(out[k - 1])
        return out

    def forward(self, d):
        out = OrderedDict()
        flag, idx = d["flag"], d["out_idx"]
        if flag == "head":                                              # resnet.py:201-251
            if idx not in (1, 2, 3, 4):
                raise AfanError(f"head out_idx must be 1..4, got {idx!r}")
            st = self.stages(d["x"], idx)
            out["low_level"], out["out"] = st[1], st[idx]
        elif flag == "tail":                                            # :253-283
            if idx not in (1, 2, 3, 4):
                raise AfanError(f"tail out_idx must be 1..4, got {idx!r}")
            out["out"] = self.run_from(d["adv"], idx)
            out["low_level"] = d["low_level_feat"]
        elif flag == "clean":                                           # :286-300
            st = self.stages(d["x"], 4)
            out["low_level"], out["out"] = st[1], st[4]
        else:
            raise AfanError(f"unknown flag {flag!r}")
        return out


def _aspp(cin, rates):
    """ASPP (1x1, three dilated 3x3, image pooling; 1x1 projection + Dropout): torchvision's module has exactly the
    reference's sub-module names (network/_deeplab.py:174-207 is the same lineage); only its Dropout rate differs."""
    from torchvision.models.segmentation.deeplabv3 import ASPP
    m = ASPP(cin, list(rates), 256)
    m.project[3].p = 0.1                                                # network/_deeplab.py:199
    return m


class DeepLabHeadV3Plus(nn.Module):
    """network/_deeplab.py:28-88: low-level 1x1 projection (48 ch), ASPP, 3x3 + 1x1 classifier on the 304-ch concat."""

    def __init__(self, cin, low_level_channels, num_classes, aspp_dilate):
        super().__init__()
        self.project = nn.Sequential(nn.Conv2d(low_level_channels, 48, 1, bias=False), nn.BatchNorm2d(48), nn.ReLU(inplace=True))
        self.aspp = _aspp(cin, aspp_dilate)
        self.classifier = nn.Sequential(nn.Conv2d(304, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                        nn.Conv2d(256, num_classes, 1))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def concat(self, low_level, aspp_out):
        low = self.project(low_level)
        up = F.interpolate(aspp_out, size=low.shape[2:], mode="bilinear", align_corners=False)
        return torch.cat([low, up], dim=1)

    def forward(self, feature, return_type=None):
        if return_type is None:
            return self.classifier(self.concat(feature["low_level"], self.aspp(feature["out"])))
        if return_type == "aspp_head":
            return self.aspp(feature["out"])
        if return_type == "aspp_tail":
            return self.classifier(self.concat(feature["low_level"], feature["adv"]))
        if return_type == "concat_head":
            return self.concat(feature["low_level"], self.aspp(feature["out"]))
        if return_type == "concat_tail":
            return self.classifier(feature["adv"])
        raise AfanError(f"unknown return_type {return_type!r}")


class SplitDeepLabV3Plus(nn.Module):
    def __init__(self, num_classes: int = 21, output_stride: int = 16, backbone: str = "resnet101", bn_momentum: float = 0.01):
        super().__init__()
        self.backbone = SplitResNet(backbone, output_stride)
        rates = [12, 24, 36] if output_stride == 8 else [6, 12, 18]
        self.classifier = DeepLabHeadV3Plus(2048, 256, num_classes, rates)
        for m in self.backbone.modules():                               # utils.set_bn_momentum(model.backbone, 0.01), main_aug_final.py:75
            if isinstance(m, nn.BatchNorm2d):
                m.momentum = bn_momentum

    def logits(self, features, size, return_type=None):
        return F.interpolate(self.classifier(features, return_type), size=size, mode="bilinear", align_corners=False)

    def forward(self, d):
        flag, idx = d["flag"], d["out_idx"]
        if flag == "head":                                              # network/utils.py:16-19
            return self.backbone(d)
        if flag not in ("tail", "clean"):
            raise AfanError(f"unknown flag {flag!r}")
        if isinstance(idx, int):                                        # :23-29
            return self.logits(self.backbone(d), d["x"].shape[-2:])
        if idx in ("aspp_head", "concat_head"):                         # :31-36
            features = self.backbone(d)
            features["adv"] = self.classifier(features, return_type=idx)
            return features
        if idx in ("aspp_tail", "concat_tail"):                         # :38-45
            return self.logits(d["adv"], d["x"].shape[-2:], return_type=idx)
        raise AfanError(f"unknown out_idx {idx!r}")


def deeplabv3plus_resnet101(num_classes=21, output_stride=8, pretrained_backbone=False):
    """Factory named like network/modeling.py:111-119 (there is no network here: pretrained_backbone must be False)."""
    if pretrained_backbone:
        raise AfanError("no network access: load a checkpoint with load_state_dict instead")
    return SplitDeepLabV3Plus(num_classes, output_stride, "resnet101")


def deeplabv3plus_resnet50(num_classes=21, output_stride=8, pretrained_backbone=False):
    if pretrained_backbone:
        raise AfanError("no network access: load a checkpoint with load_state_dict instead")
    return SplitDeepLabV3Plus(num_classes, output_stride, "resnet50")
