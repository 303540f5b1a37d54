"""CPU: host-side logic -- CLI parity with the reference flags, LR schedule, data-parallel protocol over
gloo (world_size 2), model/state-dict compatibility with the reference."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from oracle import ref_shim

PKG = importlib.import_module("cv_a-fan_b200")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_flags_and_defaults_match_reference():
    ours = PKG.main_perturb.build_parser()
    d = vars(ours.parse_args([]))
    expected = dict(steps=5, perturb_idx=13, gamma=1.5, eps=2, randinit=False, clip=False, batch_size=128, lr=0.1,
                    momentum=0.9, weight_decay=5e-4, epochs=200, decreasing_lr="50,150", print_freq=50, gpu=0,
                    seed=None, resume=False, save_dir="res56s_adv_aug", data="../data")     # main_perturb.py:28-49
    for k, v in expected.items():
        assert d[k] == v, k
    if ref_shim.available():
        ref = ref_shim.load("Classification", "main_perturb").parser
        rd = vars(ref.parse_args([]))
        for k, v in rd.items():
            assert d[k] == v, f"default of --{k} differs from the reference"
        args = ["--steps", "3", "--perturb_idx", "10", "--gamma", "0.5", "--eps", "4", "--randinit", "--clip", "--seed", "3"]
        a, b = vars(ours.parse_args(args)), vars(ref.parse_args(args))
        for k, v in b.items():
            assert a[k] == v, k


def test_lr_schedule_matches_reference():
    mpf = PKG.main_perturb
    for step in (0, 1, 175, 349, 350, 1000):
        assert mpf.warmup_lr(step, 351, 0.1) == pytest.approx(min(step * 0.1 / 350, 0.1))
    sched_model = torch.nn.Linear(1, 1)
    opt = torch.optim.SGD(sched_model.parameters(), 0.1)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[50, 150], gamma=0.1)     # main_perturb.py:75
    for epoch in range(200):
        assert mpf.multistep_lr(epoch, 0.1, [50, 150]) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-9)
        opt.step(); sch.step()


def test_model_layout_matches_reference_indexing():
    m = PKG.resnet_s.resnet56(num_classes=100)
    assert len(m.sequential_model) == 34                      # main_perturb.py:65
    assert len(PKG.resnet_s.resnet20().sequential_model) == 16   # attack_algo.py:38 default layer_number
    names = [type(l).__name__ for l in m.sequential_model]
    assert names[:4] == ["NormalizeByChannelMeanStd", "Conv2d", "DualBatchNorm2d", "ReLU"]
    assert names[4:31] == ["BasicBlock"] * 27 and names[31:] == ["AdaptiveAvgPool2d", "Flatten", "Linear"]
    assert len(m.bn_layers()) == 55
    if ref_shim.available():
        rs = ref_shim.load("Classification", "resnet_s")
        torch.manual_seed(3)
        ref = rs.ResNet(rs.BasicBlock, [9, 9, 9], num_classes=100)
        torch.manual_seed(3)
        ours = PKG.resnet_s.resnet56(num_classes=100)
        rsd, osd = ref.state_dict(), ours.state_dict()
        assert list(rsd) == list(osd)
        assert all(torch.equal(rsd[k], osd[k]) for k in rsd)          # same init stream under the same seed
        ours.load_state_dict(rsd)


def test_shard_range_and_errors():
    s = PKG.sync
    assert [s.shard_range(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    with pytest.raises(ValueError):
        s.shard_range(10, 0, 4)
    with pytest.raises(TypeError):
        s.allreduce_sums_(torch.zeros(2, 3, 2))                         # must be float64


def test_cuda_only_paths_fail_loudly_on_cpu():
    with pytest.raises(PKG.AfanError):
        PKG.attack_algo.PGD(torch.zeros(2, 3, 4, 4), lambda o, y: o.sum(), model=lambda x, **k: x, steps=1, gamma=0.1)
    with pytest.raises(PKG.AfanError):
        PKG.segmentation.mix_feature(torch.zeros(1, 4, 2, 2), torch.zeros(1, 4, 2, 2))
    with pytest.raises(PKG.AfanError):
        PKG.trainer.AfanTrainer(PKG.resnet_s.resnet20())


def _dp_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("cv_a-fan_b200")
    g = torch.Generator().manual_seed(0)
    G, N, C, H, W = 2, 8, 6, 5, 5                        # global [adv; clean] batch: G*N samples
    x = torch.randn(G * N, C, H, W, generator=g) * 1.5 + 0.4
    lo, hi = pkg.sync.shard_range(N, rank, world)
    local = torch.cat([x[gi * N + lo: gi * N + hi] for gi in range(G)])       # this rank's shard of EACH group
    n_loc = hi - lo
    xs = local.view(G, n_loc, C, H * W).double()
    sums = torch.stack([xs.sum(dim=(1, 3)), (xs * xs).sum(dim=(1, 3))], dim=-1).contiguous()    # [G, C, 2] local sums
    pkg.sync.allreduce_sums_(sums, dist.group.WORLD)                     # ONE message for both groups
    mean, var, unb, invstd = pkg.sync.stats_from_sums(sums, pkg.sync.global_count(n_loc, H * W, world), 1e-5)
    # gradient arena protocol: SUM all-reduce, 1/world folded into the update
    arena = torch.full((10,), float(rank + 1))
    scale = pkg.sync.allreduce_grad_arena_(arena, dist.group.WORLD)
    torch.save({"mean": mean, "invstd": invstd, "unb": unb, "arena": arena * scale, "x": x}, os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_protocol_reproduces_global_batch_statistics(tmp_path):
    world, port = 2, 29600 + os.getpid() % 300
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{r}.pt") for r in range(2))
    for k in ("mean", "invstd", "unb", "arena"):
        assert torch.equal(r0[k], r1[k]), k                              # every rank ends with identical values
    assert torch.allclose(r0["arena"], torch.full((10,), 1.5))           # mean of rank grads (1 and 2)
    x = r0["x"].numpy()
    rm, rv = np.zeros(6, np.float32), np.ones(6, np.float32)
    _, sm, si = orc.bn_fwd(x, np.ones(6, np.float32), np.zeros(6, np.float32), rm, rv, groups=2)    # single process, global batch
    np.testing.assert_allclose(r0["mean"].numpy(), sm, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(r0["invstd"].numpy(), si, rtol=1e-6)


def test_seg_golden_matches_the_product_model_layout():
    """tests/golden/seg_step.npz (reference execution) lists exactly the state-dict keys of cv_a-fan_b200.deeplab."""
    import numpy as np
    from oracle import seg_ref_step as ref
    g = np.load(ref.GOLDEN, allow_pickle=False)
    model = PKG.deeplab.deeplabv3plus_resnet50(num_classes=ref.NUM_CLASSES, output_stride=16)
    assert [str(k) for k in g["keys"]] == list(model.state_dict().keys())
    for name in ref.CASES:
        assert np.isfinite(g[f"{name}/losses"]).all() and g[f"{name}/losses"].shape == (ref.ITERS, 5)


def test_stat_arena_replay_equals_repeated_batchnorm_updates():
    """trainer_seg._StatArena.replay(k): the closed form new = a*after_one + b*before equals forwarding the SAME batch k
    times through train-mode BatchNorm (what the reference's repeated clean passes do, main_aug_final.py:164,166,217)."""
    torch.manual_seed(0)
    bns = [torch.nn.BatchNorm2d(5, momentum=0.01), torch.nn.BatchNorm2d(3, momentum=0.01)]
    refs = [torch.nn.BatchNorm2d(5, momentum=0.01), torch.nn.BatchNorm2d(3, momentum=0.01)]
    for b, r in zip(bns, refs):
        b.running_mean.copy_(torch.randn_like(b.running_mean)); b.running_var.copy_(torch.rand_like(b.running_var) + 0.5)
        r.load_state_dict(b.state_dict())
    arena = PKG.trainer_seg._StatArena(bns, torch.device("cpu"))
    xs = [torch.randn(4, 5, 6, 6), torch.randn(4, 3, 6, 6)]
    for k in (2, 3):
        arena.snapshot()
        for b, r, x in zip(bns, refs, xs):
            b.train()(x)                                # one pass through the arena-backed layer
            for _ in range(k):
                r.train()(x)                            # k passes through the plain layer
        arena.replay(k)
        for b, r in zip(bns, refs):
            torch.testing.assert_close(b.running_mean, r.running_mean, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(b.running_var, r.running_var, rtol=1e-5, atol=1e-6)
            assert int(b.num_batches_tracked) == int(r.num_batches_tracked)
    with pytest.raises(PKG.AfanError):
        PKG.trainer_seg._StatArena([torch.nn.BatchNorm2d(2, momentum=0.1), torch.nn.BatchNorm2d(2, momentum=0.01)], torch.device("cpu"))


def test_alias_package_shares_module_objects_with_the_real_package():
    """`python -m afan_b200.main_perturb` must see the SAME trainer / dual_bn classes as the model's
    modules -- a second copy makes every isinstance(m, DualBatchNorm2d) in the trainer fail silently."""
    import runpy
    sys.path.insert(0, ROOT)
    import afan_b200
    assert afan_b200 is PKG
    ns = runpy.run_module("afan_b200.main_perturb", run_name="not_main")        # what `python -m` executes
    assert ns["AfanTrainer"] is PKG.trainer.AfanTrainer
    assert ns["resnet_s"] is PKG.resnet_s and ns["conv"] is PKG.conv
    import afan_b200.dual_bn as alias_dual_bn
    assert alias_dual_bn.DualBatchNorm2d is PKG.dual_bn.DualBatchNorm2d
    model = ns["resnet_s"].resnet20()
    assert sum(isinstance(m, PKG.dual_bn.DualBatchNorm2d) for m in model.modules()) == 19


def test_scheduler_state_loads_into_the_reference_scheduler():
    """main_perturb.py:86 `scheduler.load_state_dict(checkpoint['scheduler'])` must accept our checkpoint."""
    mpf = PKG.main_perturb
    opt = torch.optim.SGD(torch.nn.Linear(1, 1).parameters(), 0.1)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[50, 150], gamma=0.1)
    sch.load_state_dict(mpf.scheduler_state([50, 150], 0.1, 0.1, 60, 0.01))
    assert sch.last_epoch == 60 and sch.get_last_lr() == [0.01]
    opt.step(); sch.step()
    assert sch.last_epoch == 61


def test_bench_line_stays_parseable_from_a_short_stdout_tail(tmp_path, capsys):
    """the driver keeps only the tail of stdout -- the final line must be ONE compact JSON object."""
    import json
    sys.path.insert(0, ROOT)
    import bench
    fam = {"kernel": "conv3x3 fwd/dgrad", "bound": "tensor", "achieved": 123.456789, "peak": 819.55, "unit": "TFLOP/s",
           "frac": 0.150641, "traffic": 4200000.0, "algorithmic_bytes": 4204000.0, "launches_per_step": 444,
           "avg_us": 8.123456, "share_of_step": 0.41234}
    line = {"metric": "A-FAN train img/s", "value": 12345.678901, "unit": "img/s", "n_gpus": 8, "steps": 20, "warmup": 5,
            "ms_per_step": 10.123456789, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": bench.config_dict(8), "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0,
            "reasons": ["sw_power_cap"], "samples": 31},
            "e2e": {"value": 12000.123, "unit": "img/s", "h2d_bytes_per_step": 12591104, "d2h_bytes_per_step": 32, "ms_per_step": 10.5},
            "gpu_launches": 21100, "afan_kernels_per_step": 1055, "bn_exchange": "p2p", "multi_gpu_parity": "ok",
            "multi_gpu_parity_err": {"loss_rel": 1.2e-6, "weights_abs": 3.4e-6, "ranks_bit_identical": True},
            "variants_ms_per_step": {"afan_fp32_ffma": 15.1, "afan_mma_sync_tf32": 10.4, "cudnn_fp32_nondet": 19.3, "cudnn_tf32_nondet": 12.0},
            "roofline": fam, "roofline_hbm": dict(fam, kernel="dual_bn bwd+relu", bound="hbm", unit="GB/s"),
            "reference_on_gpu_ms": {"fp32": 52.1, "tf32_default": 40.2},
            "parity_iter0": {"loss": 4.7123, "port_loss": 4.7124, "rel_err": 2e-5, "linf_max": 0.00784, "port_linf_max": 0.00784},
            "cpu_baseline": {"value": 337.0, "unit": "img/s", "cores": 16, "kind": "port",
                             "sample": "4 full iterations of this workload, batch 128, oracle port, 16 threads"}}

    class A:
        detail_file = str(tmp_path / "detail.json")
    bench.emit(line, {"kernels": [{"x": 1}] * 50}, A, 8)
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 1 and len(out[0]) <= 1400, len(out[0])
    parsed = json.loads(out[0][-1400:])
    for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "e2e", "roofline", "cpu_baseline", "clocks", "gpu_launches", "config"):
        assert k in parsed, k
    assert parsed["value"] > 0 and parsed["e2e"]["value"] > 0
    assert json.load(open(A.detail_file))["kernels"]


# ---- Detection flavour: host-side pieces of cv_a-fan_b200.faster_rcnn / trainer_det that need no GPU -------------------
def test_detection_box_arithmetic_and_losses_against_their_definitions():
    pkg = importlib.import_module("cv_a-fan_b200")
    fr = pkg.faster_rcnn
    g = torch.Generator().manual_seed(9)
    xy = torch.rand(2, 40, 2, generator=g) * 200
    src = torch.cat((xy, xy + 5 + torch.rand(2, 40, 2, generator=g) * 90), dim=2)
    xy2 = torch.rand(2, 40, 2, generator=g) * 200
    dst = torch.cat((xy2, xy2 + 5 + torch.rand(2, 40, 2, generator=g) * 90), dim=2)
    # deltas round-trip (Detection/bbox.py:42-64)
    torch.testing.assert_close(fr.apply_deltas(src, fr.box_deltas(src, dst)), dst, rtol=1e-4, atol=1e-3)
    # IoU against the scalar definition, incl. a zero-area pair -> NaN like the reference's 0/0
    iou = fr.pairwise_iou(src, dst[:, :7])
    for b, p, q in ((0, 3, 2), (1, 17, 6), (0, 39, 0)):
        a, c = src[b, p], dst[b, q]
        w = max(min(a[2], c[2]) - max(a[0], c[0]), 0.0)
        h = max(min(a[3], c[3]) - max(a[1], c[1]), 0.0)
        inter = w * h
        want = inter / ((a[2] - a[0]) * (a[3] - a[1]) + (c[2] - c[0]) * (c[3] - c[1]) - inter)
        assert abs(float(iou[b, p, q]) - float(want)) < 1e-6
    assert torch.isnan(fr.pairwise_iou(torch.zeros(1, 1, 4), torch.zeros(1, 1, 4))).all()
    assert torch.equal(fr.clip_boxes(torch.tensor([[-5.0, -1.0, 300.0, 90.0]]), 200, 80), torch.tensor([[0.0, 0.0, 200.0, 80.0]]))
    # per-image losses = the reference's loops (model.py:365-377): mean CE over the image's samples, smooth-L1 over its foreground
    s, ncls, batch = 50, 4, 3
    logits = torch.randn(s, ncls, generator=g, requires_grad=True)
    targets = torch.randint(0, ncls, (s,), generator=g)
    pred, gt = torch.randn(s, 4, generator=g, requires_grad=True), torch.randn(s, 4, generator=g)
    gt[targets == 0] = float("inf")                                   # background rows carry garbage targets in the model
    bi = torch.randint(0, batch - 1, (s,), generator=g)               # the last image gets no sample
    ce, l1 = fr.per_image_losses(logits, targets, pred, gt, bi, batch, beta=1.0)
    for i in range(batch):
        sel = (bi == i).nonzero().view(-1)
        if len(sel) == 0:
            assert torch.isnan(ce[i]) and float(l1[i]) == 0.0
            continue
        torch.testing.assert_close(ce[i], torch.nn.functional.cross_entropy(logits[sel], targets[sel]), rtol=1e-5, atol=1e-6)
        fg = sel[targets[sel] != 0]
        d = (pred[fg] - gt[fg]).abs()
        want = torch.where(d < 1.0, 0.5 * d ** 2, d - 0.5).sum() / (d.numel() + 1e-8)
        torch.testing.assert_close(l1[i], want, rtol=1e-5, atol=1e-6)
    (ce[:batch - 1].sum() + l1.sum()).backward()
    assert torch.isfinite(logits.grad).all() and torch.isfinite(pred.grad).all()       # masked rows give 0, not NaN
    assert float(pred.grad[targets == 0].abs().max()) == 0.0


def test_detection_sampler_and_anchor_invariants():
    pkg = importlib.import_module("cv_a-fan_b200")
    fr = pkg.faster_rcnn
    g = torch.Generator().manual_seed(4)
    labels = torch.randint(-1, 3, (3, 500), generator=g)
    torch.manual_seed(0)
    bi, ci = fr.select_samples(labels, fg_cap=40, total=120, sampler=fr.Sampler("reference"))
    picked = labels[bi, ci]
    assert len(bi) == 120 and int((picked > 0).sum()) == 40 and int((picked == 0).sum()) == 80 and not (picked < 0).any()
    assert len({(int(a), int(b)) for a, b in zip(bi, ci)}) == 120              # no candidate twice
    torch.manual_seed(0)                                                       # the reference's three draws, in its order
    fg, bgc = (labels > 0).nonzero(), (labels == 0).nonzero()
    fg = fg[torch.randperm(len(fg))[:40]]
    bgc = bgc[torch.randperm(len(bgc))[:80]]
    sel = torch.cat([fg, bgc])
    sel = sel[torch.randperm(len(sel))]
    assert torch.equal(sel[:, 0], bi) and torch.equal(sel[:, 1], ci)
    few = torch.tensor([[1, 0, -1, 0]])                                        # fewer candidates than asked for: take all
    bi, ci = fr.select_samples(few, fg_cap=8, total=16, sampler=fr.Sampler("reference"))
    assert sorted(ci.tolist()) == [0, 1, 3]
    rpn = fr.RegionProposalNetwork(8, [(1, 2), (1, 1), (2, 1)], [32, 64], 100, 10, 1.0, fr.Sampler("reference"))
    a = rpn.generate_anchors(image_width=96, image_height=64, num_x_anchors=3, num_y_anchors=2)
    assert a.shape == (2 * 3 * 3 * 2, 4)
    cx, cy = (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
    assert torch.allclose(cy[:18], torch.full((18,), 64 / 3)) and torch.allclose(cx[:6], torch.full((6,), 24.0))   # y major, then x
    w, h = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    torch.testing.assert_close((w * h)[:6], torch.tensor([32.0 ** 2, 64.0 ** 2] * 3), rtol=1e-5, atol=1e-2)       # ratio keeps the area
    torch.testing.assert_close((h / w)[:6], torch.tensor([0.5, 0.5, 1.0, 1.0, 2.0, 2.0]), rtol=1e-5, atol=1e-5)
    assert pkg.trainer_det.warmup_multistep_lr(0, 0.001) == pytest.approx(0.001 * 0.3333)
    assert pkg.trainer_det.warmup_multistep_lr(250, 0.001) == pytest.approx(0.001 * (0.6667 * 0.5 + 0.3333))
    assert pkg.trainer_det.warmup_multistep_lr(500, 0.001) == pytest.approx(0.001)
    assert pkg.trainer_det.warmup_multistep_lr(50000, 0.001) == pytest.approx(0.0001)
    assert pkg.trainer_det.warmup_multistep_lr(70000, 0.001) == pytest.approx(0.00001)
