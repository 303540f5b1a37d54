"""-m gpu: dual-BN forward/backward kernels vs the oracle and vs torch's own BatchNorm (the arithmetic
the reference's nn.BatchNorm2d runs).  Tolerances: 1e-5 rel / 1e-5 abs fp32 (reduction order)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as orc
from tests.util import PKG, dev

pytestmark = pytest.mark.gpu
ops = PKG.ops

CASES = [  # groups, n, c, h, w
    (1, 128, 16, 32, 32), (2, 128, 16, 32, 32), (2, 128, 32, 16, 16), (2, 128, 64, 8, 8),   # ResNet-56 tail, config 2
    (1, 4, 2048, 33, 33), (2, 2, 256, 33, 33),                                              # DeepLab-shaped, odd HW -> scalar path
    (2, 4, 256, 33, 33),                                    # plane-resident kernels at C = 256, backward stages dy / x / y (3 x 35 KB)
    (2, 2, 8, 35, 37), (1, 1, 3, 17, 19), (1, 2, 40, 129, 129),   # split path, odd H*W >= 256: peeled reduce + flat-vector / scalar apply
    (1, 3, 5, 1, 1), (2, 1, 3, 2, 2), (1, 2, 1, 1, 7), (4, 2, 6, 4, 4)]


def run_case(groups, n, c, h, w, relu, residual, replay=1, offset=0.0, split=False):
    g = torch.Generator().manual_seed(groups * 1000 + c + h)
    x = torch.randn(groups * n, c, h, w, generator=g) * 1.7 + 0.3 + offset
    res = torch.randn(x.shape, generator=g) if residual else None
    wt, b = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    dy = torch.randn(x.shape, generator=g)
    d = dev()
    rm_d, rv_d = rm.to(d), rv.to(d)
    ws = ops.bn_workspace(groups, c, d)
    y, sm, si = ops.bn_fwd(x.to(d), res.to(d) if residual else None, wt.to(d), b.to(d), rm_d, rv_d, ws,
                           groups=groups, relu=relu, replay=replay, split=split)
    dx, dres, dw, db = ops.bn_bwd(dy.to(d), x.to(d), y, wt.to(d), sm, si, ws, groups=groups, relu=relu,
                                  want_dresidual=residual, split=split)
    rm_o, rv_o = rm.numpy().copy(), rv.numpy().copy()
    y_o, sm_o, si_o = orc.bn_fwd(x.numpy(), wt.numpy(), b.numpy(), rm_o, rv_o, groups=groups,
                                 residual=res.numpy() if residual else None, relu=relu, replay=replay)
    dx_o, dres_o, dw_o, db_o = orc.bn_bwd(dy.numpy(), x.numpy(), y_o, wt.numpy(), sm_o, si_o, groups=groups,
                                          relu=relu, residual=residual)
    tol = dict(rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(y.cpu().numpy(), y_o, **tol)
    np.testing.assert_allclose(sm.cpu().numpy(), sm_o, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(si.cpu().numpy(), si_o, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rm_d.cpu().numpy(), rm_o, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv_d.cpu().numpy(), rv_o, rtol=1e-5, atol=1e-6)
    scale = max(1.0, float(np.abs(dx_o).max()))
    np.testing.assert_allclose(dx.cpu().numpy(), dx_o, rtol=1e-4, atol=2e-5 * scale)
    cnt = n * h * w
    np.testing.assert_allclose(dw.cpu().numpy(), dw_o, rtol=1e-4, atol=1e-5 * cnt ** 0.5 * groups)
    np.testing.assert_allclose(db.cpu().numpy(), db_o, rtol=1e-4, atol=1e-5 * cnt ** 0.5 * groups)
    if residual:
        # dresidual = dy masked by the ReLU of OUR y; elements whose pre-activation is within rounding of 0 may flip
        mism = dres.cpu().numpy() != dres_o
        assert mism.mean() < 1e-4
    assert int(ws.view(torch.int32)[:c].abs().sum()) == 0          # last-CTA counters are left zeroed


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("relu,residual", [(False, False), (True, False), (True, True)])
def test_bn_fwd_bwd_vs_oracle(case, relu, residual):
    run_case(*case, relu=relu, residual=residual)


@pytest.mark.parametrize("case", [(2, 128, 32, 16, 16), (2, 2, 256, 33, 33), (1, 3, 5, 1, 1), (4, 2, 6, 4, 4)])
def test_split_stats_finalize_apply_path(case):
    """The three-call form used under NCCL (stats -> all-reduce -> finalize -> apply), here with world 1."""
    run_case(*case, relu=True, residual=True, split=True)


def test_large_domain_takes_two_launch_path():
    """Per-channel domain of 32 MB: too big to stay L2-resident per cluster -> global-partials path."""
    run_case(1, 8, 4, 1024, 1024, relu=True, residual=False)


def test_replay_and_large_mean():
    run_case(2, 16, 8, 8, 8, relu=True, residual=False, replay=2)
    run_case(1, 64, 16, 16, 16, relu=False, residual=False, offset=10.0)    # |mean| >> std: E[x^2]-E[x]^2 stays accurate


def test_module_equals_two_reference_passes_through_nn_batchnorm():
    """DualBatchNorm2d(groups=2) on [adv; clean] == nn.BatchNorm2d applied to adv, then to clean
    (main_perturb.py:195-196), forward, backward, running stats and num_batches_tracked."""
    g = torch.Generator().manual_seed(5)
    n, c = 16, 32
    adv, clean = torch.randn(n, c, 8, 8, generator=g), torch.randn(n, c, 8, 8, generator=g) + 0.2
    ref = torch.nn.BatchNorm2d(c)
    ref.weight.data = torch.rand(c, generator=g) + 0.5
    ref.bias.data = torch.randn(c, generator=g)
    mod = PKG.dual_bn.DualBatchNorm2d(c)
    mod.load_state_dict(ref.state_dict())
    mod.to(dev())
    a_r, c_r = adv.clone().requires_grad_(True), clean.clone().requires_grad_(True)
    out_ref = torch.cat([F.relu(ref(a_r)), F.relu(ref(c_r))])
    dy = torch.randn(out_ref.shape, generator=g)
    out_ref.backward(dy)
    both = torch.cat([adv, clean]).to(dev()).requires_grad_(True)
    out = mod(both, relu=True, groups=2)
    out.backward(dy.to(dev()))
    torch.testing.assert_close(out.detach().cpu(), out_ref.detach(), rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(both.grad.cpu(), torch.cat([a_r.grad, c_r.grad]), rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(mod.weight.grad.cpu(), ref.weight.grad, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(mod.bias.grad.cpu(), ref.bias.grad, rtol=1e-4, atol=1e-4)
    sd = mod.state_dict()
    torch.testing.assert_close(sd["running_mean"].cpu(), ref.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sd["running_var"].cpu(), ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(sd["num_batches_tracked"]) == int(ref.num_batches_tracked) == 2
    # eval mode uses the running statistics through the affine kernel
    ref.eval(); mod.eval()
    xe = torch.randn(4, c, 8, 8, generator=g)
    torch.testing.assert_close(mod(xe.to(dev())).cpu(), ref(xe), rtol=2e-5, atol=2e-5)


def test_bn_rejects_bad_inputs():
    d = dev()
    with pytest.raises(PKG.AfanError):
        ops.bn_fwd(torch.zeros(3, 4, 2, 2, device=d), None, None, None, None, None, ops.bn_workspace(2, 4, d), groups=2)
    with pytest.raises(PKG.AfanError):      # workspace too small
        ops.bn_fwd(torch.zeros(4, 4, 2, 2, device=d), None, None, None, None, None,
                   torch.zeros(2, dtype=torch.int64, device=d), groups=2)


@pytest.mark.parametrize("case", [(2, 16, 32, 16, 16), (1, 32, 64, 8, 8), (2, 8, 64, 8, 8)])
def test_fused_p2p_exchange_loopback_two_virtual_ranks(case):
    """The fused statistics exchange (P2P stores + release/acquire flags inside the BN kernel), exercised on ONE GPU:
    two virtual ranks on two streams with mailboxes in local memory.  Each rank normalises its shard with the
    statistics of the GLOBAL batch; the oracle is single-process BN over the global batch."""
    G, n_loc, C, H, W = case
    world, d = 2, dev()
    g = torch.Generator().manual_seed(17)
    n_glob = n_loc * world
    x = torch.randn(G * n_glob, C, H, W, generator=g) * 1.3 + 0.2
    dy = torch.randn(x.shape, generator=g)
    wt, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    shard = lambda t, r: torch.cat([t[gi * n_glob + r * n_loc: gi * n_glob + (r + 1) * n_loc] for gi in range(G)]).contiguous()
    boxes = PKG.p2p.PeerMailbox.loopback(world, d, cmax=C)
    streams = [torch.cuda.Stream(device=d) for _ in range(world)]
    outs = [None] * world
    xs = [shard(x, r).to(d) for r in range(world)]
    dys = [shard(dy, r).to(d) for r in range(world)]
    rms = [torch.zeros(C, device=d) for _ in range(world)]
    rvs = [torch.ones(C, device=d) for _ in range(world)]
    wss = [ops.bn_workspace(G, C, d) for _ in range(world)]
    wt_d, b_d = wt.to(d), b.to(d)          # no host-synchronous copies while a virtual rank is waiting for its peer
    torch.cuda.synchronize()
    for r in range(world):                                  # warm each stream's allocator pool (no cudaMalloc mid-exchange)
        with torch.cuda.stream(streams[r]):
            tmp = [torch.empty_like(xs[r]) for _ in range(4)]
        del tmp
    torch.cuda.synchronize()
    for rep in range(3):                                    # several calls: exercises the sequence / ring logic
        # a virtual rank's kernel spins until its peer's kernel runs, so nothing host-synchronous (lazy module load of
        # a new kernel, pageable copies) may sit between the two launches: launch the same kernel for both ranks, then sync
        fw = [None] * world
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                fw[r] = ops.bn_fwd(xs[r], None, wt_d, b_d, rms[r], rvs[r], wss[r], groups=G, relu=True, mailbox=boxes[r])
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                y, sm, si = fw[r]
                outs[r] = (y, sm, si) + ops.bn_bwd(dys[r], xs[r], y, wt_d, sm, si, wss[r], groups=G, relu=True, mailbox=boxes[r])
        torch.cuda.synchronize()
        for bx in boxes:
            bx.check()
    assert int(boxes[0].state[0]) == 6 and int(boxes[1].state[0]) == 6          # 3 x (fwd + bwd) calls counted
    rm_o, rv_o = np.zeros(C, np.float32), np.ones(C, np.float32)
    for _ in range(3):
        y_o, sm_o, si_o = orc.bn_fwd(x.numpy(), wt.numpy(), b.numpy(), rm_o, rv_o, groups=G, relu=True)
    dx_o, _, dw_o, db_o = orc.bn_bwd(dy.numpy(), x.numpy(), y_o, wt.numpy(), sm_o, si_o, groups=G, relu=True)
    dw_sum, db_sum = 0, 0
    for r in range(world):
        y, sm, si, dx, _, dw, db = outs[r]
        np.testing.assert_allclose(y.cpu().numpy(), shard(torch.from_numpy(y_o), r).numpy(), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(sm.cpu().numpy(), sm_o, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(dx.cpu().numpy(), shard(torch.from_numpy(dx_o), r).numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(rvs[r].cpu().numpy(), rv_o, rtol=1e-5, atol=1e-6)
        dw_sum, db_sum = dw_sum + dw.cpu().numpy(), db_sum + db.cpu().numpy()
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])   # bit-identical statistics on all ranks
    np.testing.assert_allclose(dw_sum, dw_o, rtol=1e-4, atol=1e-3)              # local dweight/dbias sum to the global ones
    np.testing.assert_allclose(db_sum, db_o, rtol=1e-4, atol=1e-3)
    for bx in boxes:
        bx.close()


@pytest.mark.parametrize("shape,groups", [((2, 256, 1, 1), 1), ((4, 64, 1, 1), 2), ((2, 64, 5, 5), 1), ((4, 2048, 33, 33), 1),
                                          ((8, 256, 33, 33), 2), ((128, 32, 16, 16), 1), ((256, 64, 8, 8), 2), ((6, 8, 40, 40), 1)])
def test_statistics_survive_large_mean_over_std(shape, groups):
    """Round-2 regression (found on DeepLab's ASPP pooling branch, 2 relu-positive nearly equal values per channel): a
    sum / sum-of-squares accumulation in fp32 loses the variance when |mean| >> std.  Every forward-statistics kernel
    accumulates shifted data now; against the library's (two-pass) train-mode BatchNorm in float64."""
    g = torch.Generator().manual_seed(shape[1] + shape[2])
    x = 7.0 + 2e-3 * torch.randn(shape, generator=g)                      # mean / std = 3500
    c = shape[1]
    wt, b = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    d = dev()
    rm, rv = torch.zeros(c, device=d), torch.ones(c, device=d)
    y, sm, si = ops.bn_fwd(x.to(d), None, wt.to(d), b.to(d), rm, rv, ops.bn_workspace(groups, c, d), groups=groups)
    per = shape[0] // groups
    ref = torch.cat([F.batch_norm(x[k * per:(k + 1) * per].double(), None, None, wt.double(), b.double(), True, 0.1, 1e-5)
                     for k in range(groups)])
    # fp32 inputs carry ~6e-8 * 7 of representation noise themselves: 2e-3 of the normalised values' O(1) scale is the
    # floor for mean / std = 3500; the unshifted accumulation missed it by 0.3 .. 1.0 (variance lost entirely)
    assert float((y.double().cpu() - ref).abs().max()) < 5e-3 * float(ref.abs().max())


@pytest.mark.parametrize("shape", [(2, 8, 6, 10), (3, 5, 7, 9), (2, 64, 38, 63), (1, 256, 4, 4)])
@pytest.mark.parametrize("relu,res", [(True, True), (True, False), (False, False), (False, True)])
def test_frozen_affine_forward_backward_vs_torch(shape, relu, res):
    """afan_bn_affine_f32 / afan_bn_affine_bwd_f32 (frozen BatchNorm + residual + ReLU in one launch per direction, the
    Detection flavour's backbone) against eval-mode F.batch_norm + add + relu under autograd."""
    import torch.nn.functional as F
    d = dev()
    g = torch.Generator().manual_seed(shape[1])
    bn = torch.nn.BatchNorm2d(shape[1]).to(d).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(shape[1], generator=g) + 0.5)
        bn.bias.copy_(torch.randn(shape[1], generator=g))
        bn.running_mean.copy_(torch.randn(shape[1], generator=g))
        bn.running_var.copy_(torch.rand(shape[1], generator=g) + 0.5)
    x = torch.randn(shape, generator=g).to(d).requires_grad_(True)
    r = torch.randn(shape, generator=g).to(d).requires_grad_(True) if res else None
    dy = torch.randn(shape, generator=g).to(d)
    y = PKG.dual_bn.frozen_bn_act(x, bn, residual=r, relu=relu)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), r.grad.clone() if res else None)
    x.grad = None
    if res:
        r.grad = None
    w = bn(x) + (r if res else 0)
    w = F.relu(w) if relu else w
    w.backward(dy)
    torch.testing.assert_close(got[0], w.detach(), rtol=1e-5, atol=1e-5)
    same_mask = (got[0] > 0) == (w.detach() > 0) if relu else torch.ones_like(dy, dtype=torch.bool)
    torch.testing.assert_close(got[1][same_mask], x.grad[same_mask], rtol=1e-5, atol=1e-6)
    if res:
        torch.testing.assert_close(got[2][same_mask], r.grad[same_mask], rtol=0, atol=0)
    # the table follows the module: a write to any of the four tensors rebuilds it
    with torch.no_grad():
        bn.bias.add_(1.0)
    torch.testing.assert_close(PKG.dual_bn.frozen_bn_act(x.detach(), bn), bn(x.detach()), rtol=1e-5, atol=1e-5)
