"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE under oracle/ref_shim.py.

TEST INFRASTRUCTURE ONLY.  Run here (the container that mounts /root/reference):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

The vectors are committed; the GPU box has no /root/reference and only reads the .npz files.
Each file records inputs, every intermediate the reference exposes, and the outputs of one
reference function on the hot path.  No reference source is copied: the
functions are imported from where they lie and called.

Gradient injection: reference PGD obtains its gradient from `torch.autograd.grad(loss(model(x_adv)))`
(Classification/attack_algo.py:50-52).  To pin the UPDATE arithmetic independently of any network
we hand PGD a "model" whose output is a scalar with a prescribed gradient (including 0, -0, NaN,
+-inf, denormals) and `loss_fn = lambda out, y: out`.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
SPECIALS = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e-45, -1e-45, 1.17549435e-38, 3.4e38, -3.4e38,
                     1.0, -1.0], dtype=np.float32)


class _InjectGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, g):
        ctx.save_for_backward(g)
        return x.new_zeros(())

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return g.clone(), None


class InjectingModel:
    """Callable in all three reference model conventions; records x_adv at each call."""

    def __init__(self, grads):
        self.grads, self.calls, self.states = grads, 0, []

    def _next(self, x_adv):
        self.states.append(x_adv.detach().clone())
        g = self.grads[self.calls]
        self.calls += 1
        return _InjectGrad.apply(x_adv, g)

    def __call__(self, x, end_point=None, start_point=None):       # Classification + Segmentation
        if isinstance(x, dict):
            adv = x["adv"]
            if isinstance(adv, dict):                               # decoder_PGD: inputs['adv'] is the feature dict
                adv = adv["adv"]
            if adv is None:                                         # adv_input: the image itself is perturbed
                adv = x["x"]
            return self._next(adv)
        return self._next(x)

    def train(self):                                                # Detection: model.train().forward(...)
        return self

    def forward(self, inputs, bb, lb):
        adv = inputs["adv"]
        if isinstance(adv, dict):                                   # rpn_roi_PGD('roi')
            adv = adv["roi_output_dict"]["roi_feature_map"]
        if adv is None:                                             # Detection adv_input
            adv = inputs["x"]
        out = self._next(adv)
        z = out * 0
        if self.roi_loss_slot:                                      # only_roi_loss=True sums losses 3 and 4
            return z, z, out, z
        return out, z, z, z

    roi_loss_slot = False


def feature_like(shape, gen):
    """Post-ReLU-like activations (about half exact zeros), SURVEY 8d."""
    return torch.relu(1.5 * torch.randn(shape, generator=gen))


def grads_like(shape, steps, gen, specials=True):
    out = []
    for _ in range(steps):
        g = 1e-3 * torch.randn(shape, generator=gen)
        g[torch.rand(shape, generator=gen) < 0.05] = 0.0
        if specials:
            flat = g.view(-1)
            flat[: len(SPECIALS)] = torch.from_numpy(SPECIALS)
        out.append(g)
    return out


def gen_pgd_cases():
    cls = ref_shim.load("Classification", "attack_algo")
    seg = ref_shim.load("Segmentation", "attack_algo")
    det = ref_shim.load("Detection", "attack_algo")
    gen = torch.Generator().manual_seed(3)
    cases = {}
    shape = (4, 8, 6, 7)
    idx = 0
    for flavour in ("cls", "seg", "det"):
        for randinit in (False, True):
            for clip in (False, True):
                for gamma255, steps in ((0.5, 5), (1.5, 3)):
                    if flavour != "cls" and gamma255 == 1.5:
                        continue
                    x = feature_like(shape, gen)
                    xf = x.view(-1)
                    xf[-len(SPECIALS):] = torch.from_numpy(SPECIALS)        # special clean values too
                    grads = grads_like(shape, steps, gen)
                    gamma, eps = gamma255 / 255, 2 / 255
                    torch.manual_seed(3)
                    u = torch.rand(shape)                                   # the draw PGD will make
                    torch.manual_seed(3)
                    model = InjectingModel(grads)
                    with ref_shim.cpu_cuda_identity():
                        if flavour == "cls":
                            out = cls.PGD(x, lambda o, y: o, y=None, model=model, steps=steps, gamma=gamma,
                                          start_idx=1, layer_number=16, eps=eps, randinit=randinit, clip=clip)
                        elif flavour == "seg":
                            out = seg.PGD(x, None, None, lambda o, y: o, y=None, model=model, steps=steps,
                                          eps=eps, gamma=gamma, idx=3, randinit=randinit, clip=clip)
                        else:
                            out = det.PGD(x, None, y={"bb": None, "lb": None}, model=model, steps=steps,
                                          eps=eps, gamma=gamma, idx=3, randinit=randinit, clip=clip)
                    assert out.requires_grad and out.is_leaf
                    key = f"case{idx}"
                    idx += 1
                    cases[key + "_meta"] = np.array([gamma, eps, steps, int(randinit), int(clip),
                                                     {"cls": 0, "seg": 1, "det": 2}[flavour]], dtype=np.float64)
                    cases[key + "_x"] = x.numpy().copy()
                    cases[key + "_u"] = u.numpy().copy()
                    cases[key + "_grads"] = torch.stack(grads).numpy().copy()
                    cases[key + "_states"] = torch.stack(model.states).numpy().copy()   # x_adv BEFORE step t
                    cases[key + "_out"] = out.detach().numpy().copy()
    cases["n_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pgd_linf.npz"), **cases)
    print("pgd_linf.npz:", idx, "cases")


def gen_helper_cases():
    cls = ref_shim.load("Classification", "attack_algo")
    seg = ref_shim.load("Segmentation", "attack_algo")
    det = ref_shim.load("Detection", "attack_algo")
    gen = torch.Generator().manual_seed(5)
    out = {}
    # tensor_clamp / linfball_proj (Classification/attack_algo.py:9-19,35-36) incl. specials
    c = feature_like((3, 5, 4, 4), gen)
    t = c + 0.02 * torch.randn(c.shape, generator=gen)
    t.view(-1)[: len(SPECIALS)] = torch.from_numpy(SPECIALS)
    c.view(-1)[len(SPECIALS): 2 * len(SPECIALS)] = torch.from_numpy(SPECIALS)
    out["linf_center"], out["linf_t"] = c.numpy().copy(), t.numpy().copy()
    out["linf_radius"] = np.float32(2 / 255)
    out["linf_out"] = cls.linfball_proj(c, 2 / 255, t.clone(), in_place=True).numpy().copy()
    # l2ball_proj (attack_algo.py:21-33): inside ball, outside ball, t == center (0/0 -> NaN)
    c = feature_like((4, 3, 5, 5), gen)
    t = c.clone()
    t[0] += 1e-4 * torch.randn(t[0].shape, generator=gen)
    t[1] += 0.5 * torch.randn(t[1].shape, generator=gen)
    t[3] += 0.01 * torch.randn(t[3].shape, generator=gen)
    out["l2_center"], out["l2_t"] = c.numpy().copy(), t.numpy().copy()
    out["l2_radius"] = np.float32(0.05)
    out["l2_out"] = cls.l2ball_proj(c, 0.05, t.clone(), in_place=True).numpy().copy()
    # mix_feature (Seg :121-130, Det :254-265) and get_sample_points (Seg :108-118, Det :236-245)
    for i, shape in enumerate(((2, 16, 5, 7), (1, 256, 3, 3), (3, 2, 4, 4), (2, 40, 1, 9))):
        cl = feature_like(shape, gen)
        ad = cl + (2 / 255) * torch.sign(torch.randn(shape, generator=gen))
        if i == 0:
            cl[0, :, 0, 0] = 0.25          # constant channel column: var == 0
            ad[0, :, 0, 1] = 0.5
        out[f"mix{i}_clean"], out[f"mix{i}_adv"] = cl.numpy().copy(), ad.numpy().copy()
        out[f"mix{i}_seg"] = seg.mix_feature(cl, ad).numpy().copy()
        out[f"mix{i}_det"] = det.mix_feature(cl, ad).numpy().copy()
        for n in (3, 5):
            pts = seg.get_sample_points(cl, ad, n)
            pts_d = det.get_sample_points(cl, ad, n)
            assert all(torch.equal(a, b) for a, b in zip(pts, pts_d))
            out[f"mix{i}_pts{n}"] = torch.stack(pts).numpy().copy()
    out["n_mix"] = np.array(4)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "helpers.npz"), **out)
    print("helpers.npz written")


def gen_f1_cases():
    """decoder-side / ROI-side / input-space PGD variants of the reference (no-clip only for
    decoder_PGD and rpn_roi_PGD('roi'): their clip branches raise NameError in the reference)."""
    seg = ref_shim.load("Segmentation", "attack_algo")
    det = ref_shim.load("Detection", "attack_algo")
    gen = torch.Generator().manual_seed(13)
    out = {}

    def record(key, x, u, grads, model, result, gamma, eps, steps, randinit, clip):
        out[key + "_meta"] = np.array([gamma, eps, steps, int(randinit), int(clip)], dtype=np.float64)
        out[key + "_x"], out[key + "_u"] = x.numpy().copy(), u.numpy().copy()
        out[key + "_grads"] = torch.stack(grads).numpy().copy()
        out[key + "_out"] = result.detach().numpy().copy()

    gamma, eps = 0.4 / 255, 2 / 255
    for randinit in (False, True):
        tag = "r" if randinit else "n"
        # Seg decoder_PGD
        shape, steps = (2, 6, 5, 5), 3
        x, grads = feature_like(shape, gen), grads_like(shape, steps, gen)
        torch.manual_seed(3); u = torch.rand(shape); torch.manual_seed(3)
        model = InjectingModel(grads)
        with ref_shim.cpu_cuda_identity():
            d = seg.decoder_PGD({"adv": x.clone(), "aux": 1}, None, lambda o, y: o, y=None, model=model, steps=steps,
                                eps=eps, gamma=gamma, idx="aspp", randinit=randinit, clip=False)
        record(f"seg_decoder_{tag}", x, u, grads, model, d["adv"], gamma, eps, steps, randinit, False)
        # Seg adv_input (with clip and the final clamp to [0,1])
        shape, steps = (2, 3, 8, 8), 3
        x = torch.rand(shape, generator=gen)
        x.view(-1)[:4] = torch.tensor([0.0, 1.0, 0.001, 0.999])
        grads = grads_like(shape, steps, gen, specials=False)
        torch.manual_seed(3); u = torch.rand(shape); torch.manual_seed(3)
        model = InjectingModel(grads)
        with ref_shim.cpu_cuda_identity():
            r = seg.adv_input(x, lambda o, y: o, y=None, model=model, steps=steps, eps=eps, gamma=gamma,
                              randinit=randinit, clip=True)
        record(f"seg_advinput_{tag}", x, u, grads, model, r, gamma, eps, steps, randinit, True)
        # Det adv_input
        grads = grads_like(shape, steps, gen, specials=False)
        torch.manual_seed(3); u = torch.rand(shape); torch.manual_seed(3)
        model = InjectingModel(grads)
        with ref_shim.cpu_cuda_identity():
            r = det.adv_input(x, y={"bb": None, "lb": None}, model=model, steps=steps, eps=eps, gamma=gamma,
                              randinit=randinit, clip=True)
        record(f"det_advinput_{tag}", x, u, grads, model, r, gamma, eps, steps, randinit, True)
        # Det rpn_roi_PGD('roi'), only_roi_loss True
        shape, steps = (7, 12, 1, 1), 2
        x, grads = feature_like(shape, gen), grads_like(shape, steps, gen)
        torch.manual_seed(3); u = torch.rand(shape); torch.manual_seed(3)
        model = InjectingModel(grads)
        model.roi_loss_slot = True
        with ref_shim.cpu_cuda_identity():
            d = det.rpn_roi_PGD("roi", {"roi_output_dict": {"roi_feature_map": x.clone()}}, y={"bb": None, "lb": None},
                                model=model, steps=steps, eps=eps, gamma=gamma, randinit=randinit, clip=False)
        record(f"det_roi_{tag}", x, u, grads, model, d["roi_output_dict"]["roi_feature_map"], gamma, eps, steps, randinit, False)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "pgd_f1.npz"), **out)
    print("pgd_f1.npz:", len([k for k in out if k.endswith("_meta")]), "cases")


class _RecordingCE(nn.Module):
    def __init__(self):
        super().__init__()
        self.ce, self.values = nn.CrossEntropyLoss(), []

    def forward(self, out, y):
        v = self.ce(out, y)
        self.values.append(float(v.detach()))
        return v


def gen_train_cases():
    """Execute the unmodified reference training loop (Classification/main_perturb.py:153-225)."""
    rs = ref_shim.load("Classification", "resnet_s")
    mp = ref_shim.load("Classification", "main_perturb")
    recipes = {
        # name: (num_blocks, n_cls, perturb_idx, steps, gamma, eps, randinit, clip, batch, iters, epoch)
        "cls_train_randclip": ([1, 1, 1], 10, 6, 3, 1.0, 2.0, True, True, 4, 3, 1),
        "cls_train_shipped": ([1, 1, 1], 10, 5, 5, 0.5, 2.0, False, False, 4, 3, 1),   # cmd/run_perturb.sh:1
        "cls_train_warmup": ([2, 1, 1], 100, 5, 2, 1.5, 2.0, False, True, 4, 3, 0),     # epoch 0: warmup_lr
    }
    for name, (nb, ncls, pidx, steps, gamma, eps, randinit, clip, bs, iters, epoch) in recipes.items():
        torch.manual_seed(3)
        model = rs.ResNet(rs.BasicBlock, nb, num_classes=ncls)
        init_state = {k: v.clone() for k, v in model.state_dict().items()}
        gen = torch.Generator().manual_seed(11)
        images = [torch.rand(bs, 3, 32, 32, generator=gen) for _ in range(iters)]
        targets = [torch.randint(0, ncls, (bs,), generator=gen) for _ in range(iters)]
        with torch.no_grad():
            fshape = model(images[0], end_point=pidx, start_point=0).shape
        model.load_state_dict(init_state)          # undo the BN running-stat update of the probe
        torch.manual_seed(3)
        noises = [torch.rand(fshape) for _ in range(iters)] if randinit else []
        torch.manual_seed(3)
        mp.args = argparse.Namespace(perturb_idx=pidx, steps=steps, gamma=gamma, eps=eps, randinit=randinit,
                                     clip=clip, print_freq=10 ** 9, lr=0.1)
        mp.layer_number = len(model.sequential_model)
        crit = _RecordingCE()
        opt = torch.optim.SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=5e-4)
        with ref_shim.cpu_cuda_identity():
            top1, loss_avg, l2m, linfm = mp.train(list(zip(images, targets)), model, crit, opt, epoch)
        per_iter = np.array(crit.values, dtype=np.float64).reshape(iters, steps + 2)
        out = {"num_blocks": np.array(nb), "num_classes": np.array(ncls),
               "meta": np.array([pidx, steps, gamma, eps, int(randinit), int(clip), bs, iters, epoch],
                                dtype=np.float64),
               "images": torch.stack(images).numpy(), "targets": torch.stack(targets).numpy(),
               "ce_values": per_iter,      # per iteration: CE at each PGD step, CE(adv), CE(clean)
               "top1_avg": np.float64(top1), "loss_avg": np.float64(loss_avg),
               "l2_mean": np.float64(l2m), "linf_mean": np.float64(linfm)}
        if randinit:
            out["noises"] = torch.stack(noises).numpy()
        for k, v in init_state.items():
            out["init/" + k] = v.numpy()
        for k, v in model.state_dict().items():
            out["final/" + k] = v.detach().numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print(name, "ce:", per_iter[-1], "l2/linf mean:", float(l2m), float(linfm))


from full_case import FULL_RECIPE, full_case_inputs  # noqa: E402


def gen_train_full_case():
    """BASELINE configs[1] at FULL size: ResNet-56 / 100 classes / batch 128 / PGD-5 / perturb_idx 13 /
    rand + clip, two iterations of the unmodified reference loop (Classification/main_perturb.py:153-225).  Inputs come from
    seeds (full_case_inputs), initial weights from torch.manual_seed(3) (the product model reproduces the reference init
    stream: tests/test_host_logic.py; a checksum guards it).  Stored: every CE value, the norm means, the full small final
    tensors and every 8th element of the large ones."""
    r = FULL_RECIPE
    rs = ref_shim.load("Classification", "resnet_s")
    mp = ref_shim.load("Classification", "main_perturb")
    torch.manual_seed(r["weight_seed"])
    model = rs.ResNet(rs.BasicBlock, r["num_blocks"], num_classes=r["num_classes"])
    init_sum = float(sum(v.double().sum() for v in model.state_dict().values()))
    images, targets, noises = full_case_inputs(r)
    # the reference draws torch.rand(x.shape) from the GLOBAL CPU generator (attack_algo.py:44): seed it like noise_seed
    torch.manual_seed(r["noise_seed"])
    mp.args = argparse.Namespace(perturb_idx=r["perturb_idx"], steps=r["steps"], gamma=r["gamma"], eps=r["eps"],
                                 randinit=True, clip=True, print_freq=10 ** 9, lr=0.1)
    mp.layer_number = len(model.sequential_model)
    crit = _RecordingCE()
    opt = torch.optim.SGD(model.parameters(), 0.1, momentum=0.9, weight_decay=5e-4)
    with ref_shim.cpu_cuda_identity():
        top1, loss_avg, l2m, linfm = mp.train(list(zip(images, targets)), model, crit, opt, r["epoch"])
    per_iter = np.array(crit.values, dtype=np.float64).reshape(r["iters"], r["steps"] + 2)
    out = {"recipe": np.array(json.dumps(r)), "init_checksum": np.float64(init_sum), "ce_values": per_iter,
           "top1_avg": np.float64(top1), "loss_avg": np.float64(loss_avg), "l2_mean": np.float64(l2m),
           "linf_mean": np.float64(linfm), "noise_checksum": np.float64(sum(float(n.double().sum()) for n in noises))}
    for k, v in model.state_dict().items():
        a = v.detach().numpy()
        if a.size <= 7000:
            out["final/" + k] = a
        else:
            out["final_sub/" + k] = a.reshape(-1)[::r["sub"]].copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cls_train_full.npz"), **out)
    print("cls_train_full ce:", per_iter, "l2/linf mean:", float(l2m), float(linfm))


def gen_learnable_case():
    """Execute the unmodified reference learnable-eta loop (Classification/main_learnable.py:175-300) on a
    [3,3,3] net (16 layers) with 9 perturbation layers 4..12.  Initial weights are NOT stored: the product model
    reproduces the reference init stream under the same seed (tests/test_host_logic.py); a checksum guards it."""
    rs = ref_shim.load("Classification", "resnet_s")
    ml = ref_shim.load("Classification", "main_learnable")
    torch.manual_seed(3)
    model = rs.ResNet(rs.BasicBlock, [3, 3, 3], num_classes=10, init_weight=1 / 9)
    init_sum = float(sum(v.double().sum() for v in model.state_dict().values()))
    points, steps, gamma, eps, bs, iters = [4, 5, 6, 7, 8, 9, 10, 11, 12], 2, 1.0, 2.0, 4, 2
    gen = torch.Generator().manual_seed(17)
    images = [torch.rand(bs, 3, 32, 32, generator=gen) for _ in range(iters)]
    targets = [torch.randint(0, 10, (bs,), generator=gen) for _ in range(iters)]
    ml.args = argparse.Namespace(steps=steps, gamma=gamma, eps=eps, randinit=False, clip=True, print_freq=10 ** 9, lr=0.1,
                                 l1_coef=1.0)
    ml.perturb_idx_list, ml.layer_number = points, len(model.sequential_model)
    crit = _RecordingCE()
    opt = torch.optim.SGD(model.sequential_model.parameters(), 0.1, momentum=0.9, weight_decay=5e-4)
    opt_w = torch.optim.SGD([{"params": model.w, "lr": 0.01, "weight_decay": 0}], 0.01, momentum=0.9, weight_decay=0)
    with ref_shim.cpu_cuda_identity():
        top1, loss_avg, l2m, linfm = ml.train(list(zip(images, targets)), model, crit, opt, 1, opt_w)
    per_iter = np.array(crit.values, dtype=np.float64).reshape(iters, 9 * steps + 9 + 1)
    out = {"meta": np.array([steps, gamma, eps, bs, iters], dtype=np.float64), "points": np.array(points),
           "images": torch.stack(images).numpy(), "targets": torch.stack(targets).numpy(), "ce_values": per_iter,
           "loss_avg": np.float64(loss_avg), "l2_mean": np.asarray(l2m), "linf_mean": np.asarray(linfm),
           "init_checksum": np.float64(init_sum)}
    keep = ("w", "sequential_model.1.weight", "sequential_model.15.weight", "sequential_model.15.bias",
            "sequential_model.8.conv2.weight")
    for k, v in model.state_dict().items():
        if k in keep or "running" in k or "num_batches" in k or k.endswith("bn1.weight"):
            out["final/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cls_learnable.npz"), **out)
    print("cls_learnable.npz: loss_avg", loss_avg, "w", model.w.detach().numpy())


def gen_nms_case():
    """The reference's OWN golden vectors for NMS (Detection/test/nms/nms-large-{input,output}.npy, used by
    Detection/test/nms/test_nms.py:39-52 with threshold 0.7) repackaged as one fixture (data, not source)."""
    d = os.path.join(ref_shim.REFERENCE_ROOT, "Detection", "test", "nms")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "nms_large.npz"), input=np.load(os.path.join(d, "nms-large-input.npy")),
                        output=np.load(os.path.join(d, "nms-large-output.npy")), threshold=np.float32(0.7))
    print("nms_large.npz written")


if __name__ == "__main__":
    assert ref_shim.available(), "reference not mounted; goldens can only be generated where it is"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    which = sys.argv[1:] or ["pgd", "helpers", "train", "train_full", "f1", "learnable", "nms"]
    if "pgd" in which:
        gen_pgd_cases()
    if "helpers" in which:
        gen_helper_cases()
    if "train" in which:
        gen_train_cases()
    if "train_full" in which:
        gen_train_full_case()
    if "f1" in which:
        gen_f1_cases()
    if "learnable" in which:
        gen_learnable_case()
    if "nms" in which:
        gen_nms_case()
