"""TEST INFRASTRUCTURE ONLY -- restatement of ONE iteration of the reference's Segmentation training loop
(Segmentation/main_aug_final.py:160-232; the loop body is inline in main(), so it cannot be imported) around the
UNMODIFIED reference model (network.deeplabv3plus_*) and the UNMODIFIED Segmentation/attack_algo.py functions, both
imported under oracle/ref_shim.py.  `generate()` executes it on the CPU and writes tests/golden/seg_step.npz; the GPU
test replays the same inputs through cv_a-fan_b200.trainer_seg.SegAfanTrainer.

Only tests/ and this generator may import this file; the product package never does.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden", "seg_step.npz")

CASES = {
    # name: se, sd, steps, eps, gamma_se, gamma_sd, randinit, clip, mix_sd, noise_sd, mix_layer
    "A": dict(se=3, sd="aspp", steps=1, eps=2.0, gamma_se=0.5, gamma_sd=0.5, randinit=True, clip=False, mix_sd=True,
              noise_sd=0.5, mix_layer="01"),
    "B": dict(se=2, sd="concat", steps=2, eps=2.0, gamma_se=1.0, gamma_sd=0.5, randinit=False, clip=False, mix_sd=False,
              noise_sd=0.0, mix_layer="10"),
}
NUM_CLASSES, BATCH, SIZE, ITERS, LR, WD = 4, 2, 65, 2, 0.01, 1e-4
# small tensors stored in full (every other tensor is pinned by its L2 norm)
FULL = ("classifier.classifier.3.weight", "classifier.classifier.3.bias", "classifier.project.1.weight", "backbone.bn1.weight",
        "backbone.bn1.running_mean", "backbone.layer4.2.bn3.running_var", "classifier.aspp.project.1.running_mean",
        "backbone.layer3.5.bn2.bias", "classifier.aspp.convs.4.2.running_var")


def procedural_init(model: nn.Module, seed: int):
    """Deterministic, constructor-independent initialisation keyed on state-dict order (identical key sets in the
    reference model and in cv_a-fan_b200.deeplab)."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    with torch.no_grad():
        for k in sd:
            t = sd[k]
            if k.endswith("num_batches_tracked"):
                t.zero_()
            elif k.endswith("running_mean"):
                t.zero_()
            elif k.endswith("running_var"):
                t.fill_(1.0)
            elif k.startswith("backbone.normal."):
                continue
            elif t.dim() == 4:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                t.copy_(torch.randn(t.shape, generator=g) * (2.0 / fan_in) ** 0.5)
            elif k.endswith("weight"):                       # BatchNorm scale
                t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
            else:                                            # BatchNorm / conv bias
                t.copy_(0.05 * torch.randn(t.shape, generator=g))
    return model


def make_batches(seed: int):
    g = torch.Generator().manual_seed(seed)
    images = [torch.rand(BATCH, 3, SIZE, SIZE, generator=g) for _ in range(ITERS)]
    labels = []
    for _ in range(ITERS):
        y = torch.randint(0, NUM_CLASSES, (BATCH, SIZE, SIZE), generator=g)
        y[torch.rand(y.shape, generator=g) < 0.05] = 255     # ignore_index pixels
        labels.append(y)
    return images, labels


class TorchAttackAlgo:
    """Plain-PyTorch restatement of the Segmentation/attack_algo.py functions the iteration calls, device-agnostic and with
    the torch.rand draws injectable (`rand(shape)`), so that the SAME iteration can run on the GPU box (where
    /root/reference does not exist) as an on-device checker.  Validated against the unmodified reference functions by
    tests/test_oracle_vs_reference.py::test_seg_iteration_restatement_equals_reference (bitwise on the CPU)."""

    def __init__(self, rand=None):
        self.rand = rand if rand is not None else (lambda shape: torch.rand(shape))

    @staticmethod
    def _ascent(x_adv, loss_of, steps, gamma):
        for _ in range(steps):                                                           # attack_algo.py:49-57 / :72-82
            grad = torch.autograd.grad(loss_of(x_adv), x_adv, only_inputs=True)[0]
            x_adv.data.add_(gamma * torch.sign(grad.data))
        return x_adv

    def _start(self, x, eps, randinit):
        x_adv = x.clone()                                                                # :43-46
        if randinit:
            x_adv += (2.0 * self.rand(x_adv.shape).to(x.device) - 1.0) * eps
        return x_adv.detach().requires_grad_(True)

    def PGD(self, x, image_batch, low_level_feat, criterion, y=None, model=None, steps=3, eps=None, gamma=None, idx=1,
            randinit=False, clip=False):
        assert not clip, "the goldens do not exercise the clip branch"
        x_adv = self._start(x, eps, randinit)
        tail = lambda xa: criterion(model({"x": image_batch, "adv": xa, "out_idx": idx, "flag": "tail",
                                           "low_level_feat": low_level_feat}), y)
        return self._ascent(x_adv, tail, steps, gamma)

    def decoder_PGD(self, input_dict, image_batch, criterion, y=None, model=None, steps=3, eps=None, gamma=None, idx=1,
                    randinit=False, clip=False):
        assert not clip, "the reference's clip branch here is a NameError (attack_algo.py:81)"
        x_adv = self._start(input_dict["adv"].detach(), eps, randinit)                   # :63-69
        input_dict["adv"] = x_adv

        def tail(xa):
            return criterion(model({"x": image_batch, "adv": input_dict, "out_idx": idx + "_tail", "flag": "clean"}), y)
        self._ascent(x_adv, tail, steps, gamma)
        input_dict["adv"] = x_adv
        return input_dict

    @staticmethod
    def mix_feature(clean_feature, adv_feature):                                         # :121-130
        eps = 1e-5
        mean_cl = clean_feature.mean(dim=1, keepdim=True)
        std_cl = (clean_feature.var(dim=1, keepdim=True) + eps).sqrt()
        mean_adv = adv_feature.mean(dim=1, keepdim=True)
        std_adv = (adv_feature.var(dim=1, keepdim=True) + eps).sqrt()
        return (clean_feature - mean_cl) / std_cl * std_adv + mean_adv

    @staticmethod
    def get_sample_points(pointx, pointy, number):                                       # :108-118
        percent = 1.0 / (number - 1)
        return [pointx] + [torch.lerp(pointx, pointy, i * percent) for i in range(1, number - 1)] + [pointy]


def reference_iteration(model, attack_algo, images, labels, c, criterion, optimizer, rand=None):
    """main_aug_final.py:160-232, one pass of the `for (images, labels) in train_loader` body.  `rand` replaces the
    iteration's own torch.rand(...).cuda() draw (None: the reference expression, CPU generator)."""
    f0, f1 = int(c["mix_layer"][0]), int(c["mix_layer"][1])                                   # :26-27
    inputs_all_se = {"x": images, "adv": None, "out_idx": c["se"], "flag": "head"}             # :160
    inputs_all_sd = {"x": images, "adv": None, "out_idx": c["sd"] + "_head", "flag": "clean"}  # :161
    optimizer.zero_grad()                                                                      # :163
    output_dict_se = model(inputs_all_se)                                                      # :164
    decoder_feature_map_dict = model(inputs_all_sd)                                            # :166
    feature_map_sd = decoder_feature_map_dict["adv"].detach()                                  # :167
    low_level_feat = output_dict_se["low_level"]                                               # :169
    feature_map_se = output_dict_se["out"].detach()                                            # :170
    feature_adv_se = attack_algo.PGD(x=feature_map_se, image_batch=images, low_level_feat=low_level_feat,
                                     criterion=criterion, y=labels, model=model, steps=c["steps"], eps=c["eps"] / 255,
                                     gamma=c["gamma_se"] / 255, idx=c["se"], randinit=c["randinit"], clip=c["clip"])   # :172-184
    feature_adv_sd_dict = attack_algo.decoder_PGD(input_dict=decoder_feature_map_dict, image_batch=images,
                                                  criterion=criterion, y=labels, model=model, steps=c["steps"],
                                                  eps=c["eps"] / 255, gamma=c["gamma_sd"] / 255, idx=c["sd"],
                                                  randinit=c["randinit"], clip=c["clip"])                             # :186-197
    adv_feature_map_sd = feature_adv_sd_dict["adv"].detach()                                   # :199
    if c["mix_sd"]:
        adv_feature_map_sd = attack_algo.mix_feature(feature_map_sd, adv_feature_map_sd)       # :200-201
    if c["noise_sd"] != 0:
        u = torch.rand(adv_feature_map_sd.shape).cuda() if rand is None else rand(adv_feature_map_sd.shape)
        adv_feature_map_sd += (2.0 * u - 1.0) * c["gamma_sd"] * c["noise_sd"]                 # :202-203
    feature_adv_sd_dict["adv"] = adv_feature_map_sd                                            # :204
    adv_list_se = attack_algo.get_sample_points(feature_map_se, feature_adv_se, 3)             # :206
    if f0:
        adv_list_se[1] = attack_algo.mix_feature(feature_map_se, adv_list_se[1])               # :207-208
    if f1:
        adv_list_se[2] = attack_algo.mix_feature(feature_map_se, adv_list_se[2])               # :209-210
    clean_input_dict = {"x": images, "adv": None, "out_idx": 0, "flag": "clean"}
    adv_input_se_dict1 = {"x": images, "adv": adv_list_se[1], "out_idx": c["se"], "flag": "tail", "low_level_feat": low_level_feat}
    adv_input_se_dict2 = {"x": images, "adv": adv_list_se[2], "out_idx": c["se"], "flag": "tail", "low_level_feat": low_level_feat}
    adv_input_sd_dict = {"x": images, "adv": feature_adv_sd_dict, "out_idx": c["sd"] + "_tail", "flag": "clean"}      # :212-215
    output0, output1 = model(clean_input_dict), model(adv_input_se_dict1)                      # :217-218
    output2, output3 = model(adv_input_se_dict2), model(adv_input_sd_dict)                     # :219-220
    loss0, loss1 = criterion(output0, labels), criterion(output1, labels)                      # :222-223
    loss2, loss3 = criterion(output2, labels), criterion(output3, labels)                      # :224-225
    loss = 0.7 * loss0 + 0.1 * loss1 + 0.1 * loss2 + 0.1 * loss3                               # :229
    loss.backward()                                                                            # :231
    optimizer.step()                                                                           # :232
    return [float(loss0), float(loss1), float(loss2), float(loss3), float(loss)]


def _load_reference():
    # load the shim WITHOUT putting oracle/ on sys.path (it would shadow the `oracle` package for spawned test workers)
    try:
        from oracle import ref_shim
    except ImportError:
        import importlib.util
        spec = importlib.util.spec_from_file_location("afan_ref_shim", os.path.join(HERE, "ref_shim.py"))
        ref_shim = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_shim)
    ref_shim._install_stubs()
    tvu = types.ModuleType("torchvision.models.utils")       # removed from torchvision; network/backbone/resnet.py:3 imports it
    tvu.load_state_dict_from_url = lambda *a, **k: {}
    sys.modules.setdefault("torchvision.models.utils", tvu)
    attack_algo = ref_shim.load("Segmentation", "attack_algo")
    seg_root = os.path.join(ref_shim.REFERENCE_ROOT, "Segmentation")
    sys.path.insert(0, seg_root)
    sys.dont_write_bytecode = True
    try:
        import network
    finally:
        sys.path.remove(seg_root)
    return ref_shim, attack_algo, network


def feature_shapes(model, c, images):
    with torch.no_grad():
        sd_state = {k: v.clone() for k, v in model.state_dict().items()}
        se_shape = model({"x": images, "adv": None, "out_idx": c["se"], "flag": "head"})["out"].shape
        sd_shape = model({"x": images, "adv": None, "out_idx": c["sd"] + "_head", "flag": "clean"})["adv"].shape
        model.load_state_dict(sd_state)                      # undo the running-stat updates of the probe
    return tuple(se_shape), tuple(sd_shape)


def generate():
    ref_shim, attack_algo, network = _load_reference()
    out = {}
    for name, c in CASES.items():
        model = network.deeplabv3plus_resnet50(num_classes=NUM_CLASSES, output_stride=16, pretrained_backbone=False)
        for m in model.backbone.modules():                   # utils.set_bn_momentum(model.backbone, 0.01), :75
            if isinstance(m, nn.BatchNorm2d):
                m.momentum = 0.01
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0                                    # ASPP Dropout(0.1) would draw from an RNG the GPU cannot replay
        procedural_init(model, seed=7)
        model.train()
        images, labels = make_batches(seed=21)
        optimizer = torch.optim.SGD(params=[{"params": model.backbone.parameters(), "lr": 0.1 * LR},
                                            {"params": model.classifier.parameters(), "lr": LR}],
                                    lr=LR, momentum=0.9, weight_decay=WD)                      # :79-82
        criterion = nn.CrossEntropyLoss(ignore_index=255, reduction="mean")                    # :92
        with ref_shim.cpu_cuda_identity():
            se_shape, sd_shape = feature_shapes(model, c, images[0])
        losses = []
        for it in range(ITERS):
            torch.manual_seed(100 + it)                      # pre-draw what the iteration will draw, in its order
            if c["randinit"]:
                out[f"{name}/noise_se{it}"] = torch.rand(se_shape).numpy()
                out[f"{name}/noise_sd{it}"] = torch.rand(sd_shape).numpy()
            if c["noise_sd"] != 0:
                out[f"{name}/noise_n{it}"] = torch.rand(sd_shape).numpy()
            torch.manual_seed(100 + it)
            with ref_shim.cpu_cuda_identity():
                losses.append(reference_iteration(model, attack_algo, images[it], labels[it], c, criterion, optimizer))
            print(name, it, losses[-1], flush=True)
        out[f"{name}/losses"] = np.array(losses, dtype=np.float64)
        sd = model.state_dict()
        keys = list(sd.keys())
        out[f"{name}/norms"] = np.array([float(sd[k].double().norm()) for k in keys], dtype=np.float64)
        for k in FULL:
            out[f"{name}/final/{k}"] = sd[k].detach().numpy().copy()
    out["keys"] = np.array(keys)
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    generate()
