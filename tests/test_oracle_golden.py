"""Pin the CPU oracle (oracle/) against the golden vectors produced by the live reference
(oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import afan_ref_torch as ref_t


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bitwise(a, b, what=""):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    # NaN payloads are not part of the contract: compare NaN-ness, then bits elsewhere
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs"
    bad = (bits(a) != bits(b)) & ~na
    assert not bad.any(), f"{what}: {bad.sum()} of {a.size} elements differ bitwise"


def load_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "pgd_linf.npz"))
    for i in range(int(z["n_cases"])):
        k = f"case{i}"
        gamma, eps, steps, randinit, clip, flavour = z[k + "_meta"]
        yield dict(name=k, gamma=float(gamma), eps=float(eps), steps=int(steps), randinit=bool(randinit),
                   clip=bool(clip), flavour=int(flavour), x=z[k + "_x"], u=z[k + "_u"], grads=z[k + "_grads"],
                   states=z[k + "_states"], out=z[k + "_out"])


def test_pgd_linf_oracle_matches_reference_goldens(golden_dir):
    n = 0
    for c in load_cases(golden_dir):
        xa = orc.pgd_init_noise(c["x"], c["u"], c["eps"]) if c["randinit"] else c["x"].copy()
        for t in range(c["steps"]):
            assert_bitwise(xa, c["states"][t], f"{c['name']} state before step {t}")
            xa = orc.pgd_linf_step(c["grads"][t], c["x"] if c["clip"] else None, xa, c["gamma"], c["eps"], c["clip"])
        assert_bitwise(xa, c["out"], f"{c['name']} final")
        n += 1
    assert n == 16


def test_helpers_oracle_matches_reference_goldens(golden_dir):
    z = np.load(os.path.join(golden_dir, "helpers.npz"))
    # linfball_proj == one clip-only update (gamma = 0 keeps t, but -0 + 0 flips the sign of -0;
    # so emulate by clamp only): use the oracle step with grad = 0 and compare modulo signed zero
    out = orc.pgd_linf_step(np.zeros_like(z["linf_t"]), z["linf_center"], z["linf_t"], 0.0, float(z["linf_radius"]), True)
    ref = z["linf_out"]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    m = ~np.isnan(ref)
    assert np.array_equal(out[m], ref[m])
    # l2ball_proj: norm reduction order differs -> 2 ulp on the radius, 1e-6 rel on the result
    out = orc.l2ball_proj(z["l2_center"], float(z["l2_radius"]), z["l2_t"])
    ref = z["l2_out"]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.isnan(ref[2]).all()          # t == center -> 0/0 (attack_algo.py:29-30)
    m = ~np.isnan(ref)
    np.testing.assert_allclose(out[m], ref[m], rtol=1e-6, atol=1e-7)
    for i in range(int(z["n_mix"])):
        cl, ad = z[f"mix{i}_clean"], z[f"mix{i}_adv"]
        got = orc.mix_feature(cl, ad)
        for flavour in ("seg", "det"):
            ref = z[f"mix{i}_{flavour}"]
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6, err_msg=f"mix{i} {flavour}")
        for n in (3, 5):
            pts = orc.get_sample_points(cl, ad, n)
            ref = z[f"mix{i}_pts{n}"]
            for j in range(n):
                np.testing.assert_allclose(pts[j], ref[j], rtol=0, atol=0 if j in (0, n - 1) else 1.2e-7)


def test_delta_norms_oracle():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 3, 7, 7)).astype(np.float32)
    xa = x + (rng.standard_normal(x.shape) * 0.01).astype(np.float32)
    d, l2, linf = orc.delta_norms(xa, x)
    dt = torch.from_numpy(xa) - torch.from_numpy(x)
    assert np.array_equal(d, dt.numpy())
    np.testing.assert_allclose(l2, torch.norm(dt.reshape(5, -1), p=2, dim=1).numpy(), rtol=3e-7)
    assert np.array_equal(linf, torch.norm(dt.reshape(5, -1), p=float("inf"), dim=1).numpy())


def test_bn_oracle_matches_torch():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(6, 5, 4, 3, generator=g) * 2 + 0.5
    w, b = torch.rand(5, generator=g) + 0.5, torch.randn(5, generator=g)
    rm, rv = torch.zeros(5), torch.ones(5)
    rm_o, rv_o = rm.numpy().copy(), rv.numpy().copy()
    res = torch.randn(6, 5, 4, 3, generator=g)
    # two groups == two separate F.batch_norm passes in order (adv half first, then clean half)
    xs = x.clone().requires_grad_(True)
    ws, bs = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ys = [torch.nn.functional.batch_norm(xs[i * 3:(i + 1) * 3], rm, rv, ws, bs, True, 0.1, 1e-5) for i in range(2)]
    y_ref = torch.relu(torch.cat(ys) + res)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    y, sm, si = orc.bn_fwd(x.numpy(), w.numpy(), b.numpy(), rm_o, rv_o, groups=2, residual=res.numpy(), relu=True)
    np.testing.assert_allclose(y, y_ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rm_o, rm.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rv_o, rv.numpy(), rtol=1e-6, atol=1e-7)
    dx, dres, dw, db = orc.bn_bwd(dy.numpy(), x.numpy(), y, w.numpy(), sm, si, groups=2, relu=True, residual=True)
    np.testing.assert_allclose(dx, xs.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dw, ws.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(db, bs.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dres, (dy * (y_ref > 0)).numpy(), rtol=0, atol=0)


def test_sgd_oracle_matches_torch():
    g = torch.Generator().manual_seed(1)
    p = torch.randn(1000, generator=g).requires_grad_(True)
    opt = torch.optim.SGD([p], 0.1, momentum=0.9, weight_decay=5e-4)
    p_o, buf = p.detach().numpy().copy(), np.zeros(1000, np.float32)
    for _ in range(3):
        grad = torch.randn(1000, generator=g)
        p.grad = grad.clone()
        opt.step()
        p_o, buf = orc.sgd_momentum(p_o, grad.numpy(), buf, 0.1, 0.9, 5e-4)
    np.testing.assert_allclose(p_o, p.detach().numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["cls_train_randclip", "cls_train_shipped", "cls_train_warmup"])
def test_torch_port_replays_reference_training(golden_dir, name):
    """oracle/afan_ref_torch.py vs the unmodified reference main_perturb.train: same CPU ops in the
    same order -> losses and final weights agree to float round-off of the BLAS/oneDNN build."""
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    pidx, steps, gamma, eps, randinit, clip, bs, iters, epoch = z["meta"]
    model = ref_t.CifarResNetRef(tuple(int(v) for v in z["num_blocks"]), int(z["num_classes"]))
    model.load_state_dict({k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("init/")})
    model.train()
    opt = ref_t.make_sgd(model)
    crit = torch.nn.CrossEntropyLoss()
    iters = int(iters)
    for i in range(iters):
        if int(epoch) == 0:                       # main_perturb.py:288-293 warmup_lr
            for pg in opt.param_groups:
                pg["lr"] = min(i * 0.1 / (iters - 1), 0.1)
        noise = torch.from_numpy(z["noises"][i]) if bool(randinit) else None
        loss, out_clean, l2, linf, _ = ref_t.afan_train_iteration(
            model, opt, crit, torch.from_numpy(z["images"][i]), torch.from_numpy(z["targets"][i]),
            steps=int(steps), gamma=float(gamma), eps=float(eps), perturb_idx=int(pidx),
            randinit=bool(randinit), clip=bool(clip), noise=noise)
        ce_adv, ce_clean = z["ce_values"][i][-2:]
        np.testing.assert_allclose(float(loss), (ce_adv + ce_clean) / 2, rtol=2e-5)
    final = {k[6:]: z[k] for k in z.files if k.startswith("final/")}
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(final[k]), k
        else:
            np.testing.assert_allclose(v.numpy(), final[k], rtol=2e-3, atol=2e-5, err_msg=k)


def test_torch_port_replays_reference_training_at_full_size(golden_dir):
    """the port pinned at the BENCHMARKED scale -- ResNet-56 / 100 classes / batch 128 / PGD-5 /
    perturb_idx 13 / rand + clip, two iterations of the unmodified reference loop (tests/golden/cls_train_full.npz;
    inputs regenerated from seeds by oracle/full_case.py)."""
    import json
    from oracle.full_case import full_case_inputs
    z = np.load(os.path.join(golden_dir, "cls_train_full.npz"))
    r = json.loads(str(z["recipe"]))
    torch.manual_seed(r["weight_seed"])
    model = ref_t.CifarResNetRef(tuple(r["num_blocks"]), r["num_classes"])
    assert abs(float(sum(v.double().sum() for v in model.state_dict().values())) - float(z["init_checksum"])) < 1e-6
    images, targets, noises = full_case_inputs(r)
    assert abs(sum(float(n.double().sum()) for n in noises) - float(z["noise_checksum"])) < 1e-6
    model.train()
    opt, crit = ref_t.make_sgd(model), torch.nn.CrossEntropyLoss()
    l2s, linfs = [], []
    for i in range(r["iters"]):
        loss, _, l2, linf, _ = ref_t.afan_train_iteration(
            model, opt, crit, images[i], targets[i], steps=r["steps"], gamma=r["gamma"], eps=r["eps"],
            perturb_idx=r["perturb_idx"], randinit=True, clip=True, noise=noises[i])
        ce_adv, ce_clean = z["ce_values"][i][-2:]
        np.testing.assert_allclose(float(loss), (ce_adv + ce_clean) / 2, rtol=2e-5)
        l2s.append(l2); linfs.append(linf)
    np.testing.assert_allclose(float(torch.cat(l2s).mean()), float(z["l2_mean"]), rtol=1e-5)
    np.testing.assert_allclose(float(torch.cat(linfs).mean()), float(z["linf_mean"]), rtol=1e-6)
    sd = model.state_dict()
    for k in z.files:
        if k.startswith("final/"):
            got, ref = sd[k[6:]].numpy(), z[k]
        elif k.startswith("final_sub/"):
            got, ref = sd[k[10:]].numpy().reshape(-1)[::r["sub"]], z[k]
        else:
            continue
        if k.endswith("num_batches_tracked"):
            assert int(got) == int(ref), k
        else:
            np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-5, err_msg=k)


def test_nms_oracle_reproduces_the_references_own_golden(golden_dir):
    """Detection/test/nms/test_nms.py:39-52: 9770 boxes -> 1934 kept at threshold 0.7 (the reference's only unit test)."""
    z = np.load(os.path.join(golden_dir, "nms_large.npz"))
    det, expect = z["input"], np.sort(z["output"])
    assert det.shape == (9770, 5) and expect.shape == (1934,)
    for strict in (True, False):                     # GPU flavour (>) and CPU flavour (>=) agree on this input
        assert np.array_equal(orc.nms(det[:, :4], det[:, 4], float(z["threshold"]), strict), expect)
    # test_nms.py:19-37: empty / single / small
    assert orc.nms(np.zeros((0, 4), np.float32), np.zeros(0, np.float32), 0.7).size == 0
    assert orc.nms(np.array([[5, 5, 10, 10]], np.float32), np.array([0.8], np.float32), 0.7).tolist() == [0]
    small = np.array([[5, 5, 10, 10], [5, 5, 10, 10], [5, 5, 30, 30]], np.float32)
    assert orc.nms(small, np.array([0.6, 0.9, 0.4], np.float32), 0.7).tolist() == [1, 2]


def test_roi_align_oracle_matches_torchvision_lineage():
    """The reference's ROIAlign is the maskrcnn-benchmark kernel; torchvision.ops.roi_align(aligned=False) on CPU is its
    direct descendant (third-party check of the restatement, torchvision 0.26)."""
    tv_ops = pytest.importorskip("torchvision.ops")
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(2, 5, 19, 23, generator=g)
    rois = torch.tensor([[0, 3.0, 5.0, 120.0, 90.0], [1, 0.0, 0.0, 40.0, 30.0], [0, 200.0, 10.0, 400.0, 300.0],
                         [1, 50.0, 60.0, 50.5, 60.5], [0, -20.0, -10.0, 30.0, 40.0]])
    for ratio in (0, 2):
        ft = feat.clone().requires_grad_(True)
        ref = tv_ops.roi_align(ft, rois, (7, 7), spatial_scale=1 / 16, sampling_ratio=ratio if ratio else -1, aligned=False)
        dy = torch.randn(ref.shape, generator=g)
        ref.backward(dy)
        out, dfeat = orc.roi_align(feat.numpy(), rois.numpy(), (7, 7), 1 / 16, ratio, dout=dy.numpy())
        np.testing.assert_allclose(out, ref.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dfeat, ft.grad.numpy(), rtol=1e-4, atol=1e-5)
