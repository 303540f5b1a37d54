"""CPU, only where /root/reference is mounted (this container): fresh random cross-checks of the oracle
restatement against the LIVE, unmodified reference functions (beyond the committed goldens)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import afan_ref_torch as ref_t
from oracle import ref_shim
from oracle.gen_golden import InjectingModel, feature_like, grads_like

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("randinit,clip", [(False, False), (True, True), (False, True)])
def test_pgd_update_oracle_vs_live_reference(seed, randinit, clip):
    cls = ref_shim.load("Classification", "attack_algo")
    gen = torch.Generator().manual_seed(seed)
    shape, steps, gamma, eps = (3, 7, 5, 6), 4, 1.5 / 255, 2 / 255
    x, grads = feature_like(shape, gen), grads_like(shape, steps, gen)
    torch.manual_seed(seed)
    u = torch.rand(shape)
    torch.manual_seed(seed)
    with ref_shim.cpu_cuda_identity():
        out = cls.PGD(x, lambda o, y: o, model=InjectingModel(grads), steps=steps, gamma=gamma, eps=eps,
                      randinit=randinit, clip=clip).detach().numpy()
    xa = orc.pgd_init_noise(x.numpy(), u.numpy(), eps) if randinit else x.numpy().copy()
    for t in range(steps):
        xa = orc.pgd_linf_step(grads[t].numpy(), x.numpy() if clip else None, xa, gamma, eps, clip)
    nan = np.isnan(out)
    assert np.array_equal(nan, np.isnan(xa))
    assert np.array_equal(out[~nan].view(np.uint32), xa[~nan].view(np.uint32))
    # the torch port used as the CPU baseline is the same computation
    port = ref_t.pgd_reference(x, lambda o, y: o, None, InjectingModel(grads), steps, gamma, 1, 16, eps, randinit, clip,
                               noise=u).detach().numpy()
    assert np.array_equal(port[~nan].view(np.uint32), out[~nan].view(np.uint32))


def test_mix_and_l2_oracle_vs_live_reference():
    seg = ref_shim.load("Segmentation", "attack_algo")
    cls = ref_shim.load("Classification", "attack_algo")
    gen = torch.Generator().manual_seed(9)
    cl = feature_like((2, 48, 9, 11), gen)
    ad = cl + 0.01 * torch.randn(cl.shape, generator=gen)
    np.testing.assert_allclose(orc.mix_feature(cl.numpy(), ad.numpy()), seg.mix_feature(cl, ad).numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ref_t.mix_feature_reference(cl, ad).numpy(), seg.mix_feature(cl, ad).numpy(), rtol=0, atol=0)
    t = cl + 0.1 * torch.randn(cl.shape, generator=gen)
    ref = cls.l2ball_proj(cl, 0.5, t.clone()).numpy()
    np.testing.assert_allclose(orc.l2ball_proj(cl.numpy(), 0.5, t.numpy()), ref, rtol=1e-6, atol=1e-7)


def test_seg_iteration_restatement_equals_reference():
    """oracle/seg_ref_step.TorchAttackAlgo + cv_a-fan_b200.deeplab on the CPU reproduce the golden losses that the
    UNMODIFIED reference model + attack_algo produced (tests/golden/seg_step.npz): pins the on-device checker of
    tests/test_gpu_seg.py and the product model in one go."""
    import importlib
    import numpy as np
    from oracle import seg_ref_step as ref
    pkg = importlib.import_module("cv_a-fan_b200")
    g = np.load(ref.GOLDEN, allow_pickle=False)
    for name, c in ref.CASES.items():
        model = pkg.deeplab.deeplabv3plus_resnet50(num_classes=ref.NUM_CLASSES, output_stride=16)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        ref.procedural_init(model, seed=7).train()
        images, labels = ref.make_batches(seed=21)
        opt = torch.optim.SGD(params=[{"params": model.backbone.parameters(), "lr": 0.1 * ref.LR},
                                      {"params": model.classifier.parameters(), "lr": ref.LR}],
                              lr=ref.LR, momentum=0.9, weight_decay=ref.WD)
        crit = torch.nn.CrossEntropyLoss(ignore_index=255, reduction="mean")
        it = 0                                            # one iteration is enough to pin every branch; keeps the CPU suite short
        draws = []
        if c["randinit"]:
            draws += [torch.from_numpy(g[f"{name}/noise_se{it}"]), torch.from_numpy(g[f"{name}/noise_sd{it}"])]
        if c["noise_sd"] != 0:
            draws.append(torch.from_numpy(g[f"{name}/noise_n{it}"]))
        q = iter(draws)
        rand = lambda shape: next(q)
        losses = ref.reference_iteration(model, ref.TorchAttackAlgo(rand), images[it], labels[it], c, crit, opt, rand=rand)
        np.testing.assert_allclose(losses, g[f"{name}/losses"][it], rtol=1e-6)


def test_split_deeplab_equals_reference_model_in_every_protocol_mode():
    """cv_a-fan_b200.deeplab.SplitDeepLabV3Plus loads the reference model's state dict and reproduces it bit for bit in
    all 13 modes of the dict protocol (network/utils.py:14-46, backbone/resnet.py:198-304, _deeplab.py:46-80)."""
    import importlib
    from oracle import seg_ref_step as ref
    pkg = importlib.import_module("cv_a-fan_b200")
    shim, _, network = ref._load_reference()
    torch.manual_seed(0)
    theirs = network.deeplabv3plus_resnet50(num_classes=5, output_stride=16, pretrained_backbone=False)
    mine = pkg.deeplab.deeplabv3plus_resnet50(num_classes=5, output_stride=16)
    assert list(mine.state_dict().keys()) == list(theirs.state_dict().keys())
    mine.load_state_dict(theirs.state_dict())
    for m in (theirs, mine):
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    x = torch.rand(2, 3, 65, 65)
    with shim.cpu_cuda_identity(), torch.no_grad():
        for idx in (1, 2, 3, 4):
            r = theirs({"x": x, "adv": None, "out_idx": idx, "flag": "head"})
            m = mine({"x": x, "adv": None, "out_idx": idx, "flag": "head"})
            assert torch.equal(r["out"], m["out"]) and torch.equal(r["low_level"], m["low_level"])
            rt = theirs({"x": x, "adv": r["out"], "out_idx": idx, "flag": "tail", "low_level_feat": r["low_level"]})
            mt = mine({"x": x, "adv": m["out"], "out_idx": idx, "flag": "tail", "low_level_feat": m["low_level"]})
            assert torch.equal(rt, mt), idx
        assert torch.equal(theirs({"x": x, "adv": None, "out_idx": 0, "flag": "clean"}),
                           mine({"x": x, "adv": None, "out_idx": 0, "flag": "clean"}))
        for sd in ("aspp", "concat"):
            r = theirs({"x": x, "adv": None, "out_idx": sd + "_head", "flag": "clean"})
            m = mine({"x": x, "adv": None, "out_idx": sd + "_head", "flag": "clean"})
            assert torch.equal(r["adv"], m["adv"]), sd
            assert torch.equal(theirs({"x": x, "adv": r, "out_idx": sd + "_tail", "flag": "clean"}),
                               mine({"x": x, "adv": m, "out_idx": sd + "_tail", "flag": "clean"})), sd


def _det_models(monkeypatch):
    """(reference Detection model under the shim, cv_a-fan_b200.faster_rcnn.FasterRCNN) with identical weights.  On this
    CPU-only box the package's two native entry points (NMS, ROIAlign) are replaced by the C oracle -- the host logic under
    test is everything else (labels, sampling, losses, split protocol); the kernels have their own GPU parity tests."""
    import importlib
    from oracle import det_ref_step as D
    ref_shim_, attack_algo, det_model, backbone_base, r101, Pooler = D._load_reference()
    pkg = importlib.import_module("cv_a-fan_b200")
    frcnn, detection = pkg.faster_rcnn, pkg.detection
    nms_ref, roi_ref = D.install_support_stubs(orc)

    def nms_flags(boxes, scores, thr):
        keep = torch.zeros(boxes.shape[0], dtype=torch.uint8)
        idx = nms_ref(boxes, scores, thr)
        keep[idx] = 1
        return keep, torch.tensor([idx.numel()], dtype=torch.int32)

    def nms_batched(boxes_sorted, thr, max_keep, want_flags=False):
        b, n = boxes_sorted.shape[:2]
        kept, counts = torch.zeros(b, max_keep, 4), torch.zeros(b, dtype=torch.int32)
        for i in range(b):
            idx = nms_ref(boxes_sorted[i], -torch.arange(n, dtype=torch.float32), thr)[:max_keep]
            kept[i, :idx.numel()] = boxes_sorted[i][idx]
            counts[i] = idx.numel()
        return kept, counts, None

    monkeypatch.setattr(pkg.ops, "nms_flags", nms_flags)
    monkeypatch.setattr(pkg.ops, "nms_batched", nms_batched)
    monkeypatch.setattr(detection, "roi_align", roi_ref)
    with ref_shim_.cpu_cuda_identity():
        ref = D.build_reference_model(det_model, backbone_base, r101, Pooler)
    m = D.MODEL
    ours = frcnn.FasterRCNN(m["num_classes"], anchor_ratios=m["anchor_ratios"], anchor_sizes=m["anchor_sizes"],
                            rpn_pre_nms_top_n=m["pre_nms"], rpn_post_nms_top_n=m["post_nms"], layers=m["layers"],
                            base_width=m["base_width"], sampler="reference")
    D.procedural_init(ours, 7)
    return D, ref_shim_, attack_algo, ref, ours


def test_split_faster_rcnn_state_dict_is_a_subset_of_the_reference(monkeypatch):
    D, _, _, ref, ours = _det_models(monkeypatch)
    rsd, osd = ref.state_dict(), ours.state_dict()
    assert set(osd) <= set(rsd)
    extra = set(rsd) - set(osd)
    assert all(k.startswith("_bn_modules.") for k in extra), sorted(extra)[:5]      # aliases of BatchNorm tensors only
    for k in osd:
        assert osd[k].shape == rsd[k].shape, k
        if not k.startswith("features.normal."):
            assert torch.equal(osd[k], rsd[k]), k                                       # the key-keyed initialisation agrees
    # requires_grad pattern (frozen BatchNorm + conv1 / layer1)
    rg = {k: p.requires_grad for k, p in ref.named_parameters()}
    for k, p in ours.named_parameters():
        assert p.requires_grad == rg[k], k
    assert ours.load_reference_state_dict(rsd) == len(osd)


def test_split_faster_rcnn_equals_reference_model_in_every_protocol_mode(monkeypatch):
    D, shim, attack_algo, ref, ours = _det_models(monkeypatch)
    images, boxes, classes = D.make_batches(33)
    x, bb, lb = images[0], boxes[0], classes[0]
    close = lambda a, b: torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-6)

    def both(d_ref, d_ours, seed=11):
        torch.manual_seed(seed)
        with shim.cpu_cuda_identity():
            r = ref.train().forward(d_ref, bb, lb)
        torch.manual_seed(seed)
        o = ours.train()(d_ours, bb, lb)
        return r, o

    for se in (1, 2, 3):
        r, o = both({"x": x, "adv": None, "out_idx": se, "flag": "head"}, {"x": x, "adv": None, "out_idx": se, "flag": "head"})
        close(o, r)
        feat = r.detach()
        r4, o4 = both({"x": x, "adv": feat, "out_idx": se, "flag": "tail"}, {"x": x, "adv": feat.clone(), "out_idx": se, "flag": "tail"})
        for a, b in zip(o4, r4):
            close(a, b)
    r4, o4 = both({"x": x, "adv": None, "out_idx": 0, "flag": "clean"}, {"x": x, "adv": None, "out_idx": 0, "flag": "clean"})
    for a, b in zip(o4, r4):
        close(a, b)
    assert float(r4[3][0]) > 0 and float(r4[3][1]) == 0          # one image with foreground proposals, one without
    # gradients of the summed losses w.r.t. a trainable backbone weight, the RPN and the ROI head
    names = ("features.layer2.0.conv1.weight", "features.layer3.1.conv2.weight", "rpn._features.0.weight",
             "rpn._anchor_transformer.bias", "detection._proposal_transformer.weight", "features.layer4.0.conv2.weight")
    gr = torch.autograd.grad(attack_algo.compute_loss(*r4), [dict(ref.named_parameters())[n] for n in names])
    go = torch.autograd.grad(sum(t.mean() for t in o4), [dict(ours.named_parameters())[n] for n in names])
    for a, b in zip(go, gr):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    # ROI-side split: 'roi_head' dictionary, then 'roi_tail' on it
    rd, od = both({"x": x, "adv": None, "out_idx": "roi_head", "flag": "clean"}, {"x": x, "adv": None, "out_idx": "roi_head", "flag": "clean"})
    close(od["roi_output_dict"]["roi_feature_map"], rd["roi_output_dict"]["roi_feature_map"])
    assert torch.equal(od["roi_output_dict"]["batch_indices"], rd["roi_output_dict"]["batch_indices"])
    assert torch.equal(od["roi_output_dict"]["gt_proposal_classes"], rd["roi_output_dict"]["gt_proposal_classes"])
    fgm = rd["roi_output_dict"]["gt_proposal_classes"] > 0
    close(od["roi_output_dict"]["gt_proposal_transformers"][fgm], rd["roi_output_dict"]["gt_proposal_transformers"][fgm])
    r4, o4 = both({"adv": rd, "out_idx": "roi_tail", "flag": "clean"}, {"adv": od, "out_idx": "roi_tail", "flag": "clean"})
    for a, b in zip(o4, r4):
        close(a, b)
    # RPN-side split
    rd, od = both({"x": x, "adv": None, "out_idx": "rpn_head", "flag": "clean"}, {"x": x, "adv": None, "out_idx": "rpn_head", "flag": "clean"})
    close(od["rpn_feature_map_dict"]["rpn_feature"], rd["rpn_feature_map_dict"]["rpn_feature"])
    close(od["anchor_bboxes"], rd["anchor_bboxes"])
    # 'rpn_tail' (model.py:98-113) is the one branch that does not re-freeze BatchNorm after the caller's model.train(): in
    # the reference layer4's BatchNorm then runs on batch statistics there.  The branch is unreachable from the training
    # iteration (train_aug_final.py:88 raises KeyError for pertub_idx_sd='rpn'); this package keeps BatchNorm frozen in
    # every branch, so the comparison freezes the reference's by hand.
    torch.manual_seed(11)
    with shim.cpu_cuda_identity():
        ref.train()
        for bn in ref._bn_modules:
            bn.eval()
        r4 = ref.forward({"adv": rd, "out_idx": "rpn_tail", "flag": "clean"}, bb, lb)
    torch.manual_seed(11)
    o4 = ours.train()({"adv": od, "out_idx": "rpn_tail", "flag": "clean"}, bb, lb)
    for a, b in zip(o4, r4):
        close(a, b)
    # inference protocol
    with torch.no_grad():
        with shim.cpu_cuda_identity():
            r = ref.eval().forward({"x": x, "adv": None, "out_idx": 0, "flag": "clean"})
        o = ours.eval()({"x": x, "adv": None, "out_idx": 0, "flag": "clean"})
    assert r[0].shape == o[0].shape
    close(o[0], r[0])
    assert torch.equal(o[1], r[1]) and torch.equal(o[3], r[3])
    close(o[2], r[2])
