"""afan_b200 -- B200-native (sm_100a) implementation of A-FAN's adversarial-feature inner loop.

The directory is named `cv_a-fan_b200` (not a valid identifier): import it with
`importlib.import_module("cv_a-fan_b200")` or through the `afan_b200` alias module at the repo root.

Public surface (mirrors the reference's Python interface for this path):
    attack_algo        PGD / linfball_proj / l2ball_proj / tensor_clamp      (Classification/attack_algo.py)
    segmentation       PGD / mix_feature / get_sample_points                 (Segmentation/attack_algo.py)
    detection          PGD / mix_feature / get_sample_points                 (Detection/attack_algo.py)
    dual_bn            DualBatchNorm2d (clean/adversarial statistics in one sweep)
    resnet_s           splittable CIFAR ResNet (Classification/resnet_s.py)
    faster_rcnn        splittable Faster R-CNN R101-C4 (Detection/model.py) + trainer_det.DetAfanTrainer (train_aug_final.py:78-163)
    trainer            AfanTrainer: head-cached, CUDA-graphed A-FAN training step (main_perturb.py:173-201)
    ops, _lib          tensor-level / ctypes bindings of include/afan_b200.h
"""
from . import _lib, ops  # noqa: F401
from ._lib import AfanError, version  # noqa: F401
from . import attack_algo, conv, deeplab, detection, dual_bn, faster_rcnn, main_perturb, p2p, prefetch, resnet_s, segmentation, sync, trainer, trainer_det, trainer_learnable, trainer_seg  # noqa: F401,E402

__all__ = ["ops", "attack_algo", "segmentation", "detection", "dual_bn", "resnet_s", "trainer", "AfanError", "version"]
