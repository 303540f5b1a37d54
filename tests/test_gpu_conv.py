"""Hand-written 3x3 convolutions (forward / input gradient / weight gradient) against an fp64 convolution of the same
inputs -- the tail's nn.Conv2d of Classification/resnet_s.py:53,55.  Tolerance: fp32 accumulation over K = 9*C terms,
|err| <= 2e-5 * max|ref| (cuDNN's own fp32 algorithms measure 1e-5 .. 5e-5 on these shapes)."""
import importlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

pkg = importlib.import_module("cv_a-fan_b200")
ops, conv = pkg.ops, pkg.conv

SHAPES = [(128, 16, 32), (128, 32, 16), (128, 64, 8), (256, 32, 16), (3, 64, 32), (5, 32, 8), (2, 16, 8), (7, 16, 16),
          (1, 32, 32), (9, 64, 16)]


def _rel(a, ref):
    return ((a.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("n,c,h", SHAPES)
def test_conv3x3_forward_dgrad_wgrad_vs_fp64(n, c, h, monkeypatch):
    monkeypatch.setattr(conv, "MODE", "afan")            # the strict-fp32 FFMA kernels
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(n * 1000 + c + h)
    x = torch.randn(n, c, h, h, generator=g).to(dev)
    w = (torch.randn(c, c, 3, 3, generator=g) * (2.0 / (9 * c)) ** 0.5).to(dev)
    dy = torch.randn(n, c, h, h, generator=g).to(dev)
    m = conv.Conv3x3(c, c, 1).to(dev)
    with torch.no_grad():
        m.weight.copy_(w)
    xr = x.clone().requires_grad_(True)
    y = m(xr)
    dx, dw = torch.autograd.grad(y, (xr, m.weight), dy)
    ref = F.conv2d(x.double(), w.double(), padding=1)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), padding=1)
    ref_dw = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), padding=1)
    assert _rel(y, ref) < 2e-5 and _rel(dx, ref_dx) < 2e-5 and _rel(dw, ref_dw) < 2e-5
    # zero padding really is zero: a constant image convolved with an all-ones kernel counts the taps inside the image
    with torch.no_grad():
        m.weight.fill_(1.0)
    ones = m(torch.ones(1, c, h, h, device=dev))
    assert ones[0, 0, 0, 0].item() == 4 * c and ones[0, 0, 1, 1].item() == 9 * c and ones[0, -1, -1, 0].item() == 4 * c
    assert ones[0, 0, 0, 1].item() == 6 * c


@pytest.mark.parametrize("mode", ["afan", "tc3"])
def test_conv3x3_is_deterministic_and_repacks_after_weight_update(mode, monkeypatch):
    monkeypatch.setattr(conv, "MODE", mode)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = conv.Conv3x3(32, 32, 1).to(dev)
    x = torch.randn(16, 32, 16, 16, device=dev, requires_grad=True)
    dy = torch.randn(16, 32, 16, 16, device=dev)
    outs = [torch.autograd.grad(m(x), (x, m.weight), dy) for _ in range(3)]
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
    y0 = m(x).detach()
    with torch.no_grad():
        m.weight.mul_(2.0)             # bumps weight._version -> the module repacks
    assert torch.allclose(m(x).detach(), 2 * y0, rtol=1e-6, atol=1e-6)


def test_unsupported_shapes_use_the_library_convolution():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = conv.Conv3x3(16, 32, 2).to(dev)                      # a stride-2 convolution on a map size the kernels do not cover
    x = torch.randn(4, 16, 24, 24, device=dev)
    assert torch.allclose(m(x), F.conv2d(x, m.weight, stride=2, padding=1))
    m2 = conv.Conv3x3(16, 16, 1).to(dev)
    x2 = torch.randn(2, 16, 12, 12, device=dev)              # 12x12 is not a covered map size
    assert torch.allclose(m2(x2), F.conv2d(x2, m2.weight, padding=1))
    assert ops.conv3x3_supported(torch.randn(2, 16, 16, 16, device=dev), m2.weight)


@pytest.mark.parametrize("mode,tol", [("3xtf32", 6e-5), ("tf32", 5e-3)])
@pytest.mark.parametrize("n,c,h", [(16, 16, 32), (16, 32, 16), (16, 64, 8), (3, 64, 32), (5, 32, 8), (2, 16, 8)])
def test_tensor_core_twins_vs_fp64(mode, tol, n, c, h, monkeypatch):
    """mma.sync TF32 kernels: 3xTF32 split within ~1e-5 of fp64 (cuDNN's fp32 algorithms: 1e-5 .. 5e-5), one-pass TF32
    within TF32's 2^-11 input rounding (opt-in mode)."""
    monkeypatch.setattr(conv, "MODE", mode)
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(n + c + h)
    x = torch.randn(n, c, h, h, generator=g).to(dev).requires_grad_(True)
    dy = torch.randn(n, c, h, h, generator=g).to(dev)
    m = conv.Conv3x3(c, c, 1).to(dev)
    y = m(x)
    dx, dw = torch.autograd.grad(y, (x, m.weight), dy)
    w = m.weight.detach()
    ref = F.conv2d(x.detach().double(), w.double(), padding=1)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), padding=1)
    ref_dw = torch.nn.grad.conv2d_weight(x.detach().double(), w.shape, dy.double(), padding=1)
    assert _rel(y, ref) < tol and _rel(dx, ref_dx) < tol
    assert _rel(dw, ref_dw) < 2e-5                       # the weight gradient stays on the strict-fp32 kernel


def test_identity_shortcut_gradient_is_added_in_the_dgrad_epilogue(monkeypatch):
    """BasicBlock (resnet_s.py:45-77): with the hand-written kernels the gradient of `out += shortcut(x)` joins conv1's
    input gradient inside the dgrad kernel (`addend`); the block's gradients must equal the library-convolution block's."""
    dev = torch.device("cuda:0")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        blk = pkg.resnet_s.BasicBlock(32, 32, 1).to(dev).train()
        x0 = torch.randn(8, 32, 16, 16, device=dev)
        dy = torch.randn(8, 32, 16, 16, device=dev)
        res = {}
        for mode in ("afan", "cudnn"):
            monkeypatch.setattr(conv, "MODE", mode)
            for m in blk.modules():
                if isinstance(m, pkg.dual_bn.DualBatchNorm2d):
                    m.running_mean.zero_(); m.running_var.fill_(1.0)
            x = x0.clone().requires_grad_(True)
            y = blk(x)
            grads = torch.autograd.grad(y, (x, blk.conv1.weight, blk.conv2.weight, blk.bn1.weight, blk.bn2.bias), dy)
            res[mode] = (y.detach(),) + grads
        for a, b in zip(res["afan"], res["cudnn"]):
            torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-4)
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("n,cin,hin", [(128, 16, 32), (128, 32, 16), (3, 16, 32), (5, 32, 16), (256, 32, 16)])
def test_stride2_transition_forward_and_dgrad_vs_fp64(n, cin, hin, monkeypatch):
    """Stage transitions (resnet_s.py:98): C -> 2C, stride 2; forward, input gradient and weight gradient hand-written."""
    dev = torch.device("cuda:0")
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)           # the library wgrad is compared at fp32 accuracy
    g = torch.Generator(device="cpu").manual_seed(n + cin)
    x = torch.randn(n, cin, hin, hin, generator=g).to(dev).requires_grad_(True)
    dy = torch.randn(n, 2 * cin, hin // 2, hin // 2, generator=g).to(dev)
    m = conv.Conv3x3(cin, 2 * cin, 2).to(dev)
    y = m(x)
    assert y.shape == (n, 2 * cin, hin // 2, hin // 2)
    dx, dw = torch.autograd.grad(y, (x, m.weight), dy)
    w = m.weight.detach().double()
    ref = F.conv2d(x.detach().double(), w, stride=2, padding=1)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w, dy.double(), stride=2, padding=1)
    ref_dw = torch.nn.grad.conv2d_weight(x.detach().double(), w.shape, dy.double(), stride=2, padding=1)
    assert _rel(y, ref) < 2e-5 and _rel(dx, ref_dx) < 2e-5 and _rel(dw, ref_dw) < 2e-5
    outs = [torch.autograd.grad(m(x), (x, m.weight), dy) for _ in range(2)]
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], dx)          # deterministic
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][1], dw)


@pytest.mark.parametrize("n,c,h", [(128, 32, 16), (3, 32, 16), (1, 32, 16), (128, 64, 8), (6, 64, 8), (256, 64, 8), (256, 32, 16)])
def test_tcgen05_conv_vs_fp64(n, c, h, monkeypatch):
    """tcgen05 (UMMA) implicit-GEMM convolution, 3xTF32 split, TMEM accumulators: forward, input gradient and weight
    gradient within 6e-5 of an fp64 convolution (the FFMA kernel and cuDNN's fp32 algorithms measure 1e-5 .. 5e-5)."""
    monkeypatch.setattr(conv, "MODE", "tc3")
    dev = torch.device("cuda:0")
    assert ops.conv3x3_umma_supported(n, c, h)
    g = torch.Generator(device="cpu").manual_seed(n * 1000 + c + h)
    x = torch.randn(n, c, h, h, generator=g).to(dev)
    w = (torch.randn(c, c, 3, 3, generator=g) * (2.0 / (9 * c)) ** 0.5).to(dev)
    dy = torch.randn(n, c, h, h, generator=g).to(dev)
    m = conv.Conv3x3(c, c, 1).to(dev)
    with torch.no_grad():
        m.weight.copy_(w)
    xr = x.clone().requires_grad_(True)
    y, tap = m.forward_with_tap(xr)                      # identity-shortcut flavour: dgrad epilogue adds the tap gradient
    dtap = torch.randn(n, c, h, h, generator=g).to(dev)
    dx, dw = torch.autograd.grad((y, tap), (xr, m.weight), (dy, dtap))
    ref = F.conv2d(x.double(), w.double(), padding=1)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), padding=1) + dtap.double()
    ref_dw = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), padding=1)
    assert _rel(y, ref) < 6e-5, _rel(y, ref)
    assert _rel(dx, ref_dx) < 6e-5, _rel(dx, ref_dx)
    assert _rel(dw, ref_dw) < 6e-5, _rel(dw, ref_dw)           # C = 32: tcgen05 weight gradient (3xTF32, pixels as the reduction dim)
    dw2 = torch.autograd.grad(m.forward_with_tap(xr)[0], m.weight, dy)[0]
    assert torch.equal(dw, dw2)                                # fixed-order fold: bitwise reproducible
    # both weight-gradient kernels directly (the module picks the tcgen05 one for C = 32 only, where it measured faster)
    ws = ops.conv3x3_wgrad_workspace(c, dev)
    for math in ("fp32", "umma"):
        assert _rel(ops.conv3x3_wgrad(x, dy, ws, math=math), ref_dw) < 6e-5, math
    acc = torch.ones(c, c, 3, 3, device=dev)
    ops.conv3x3_wgrad(x, dy, ws, accumulate_into=acc, math="umma")
    assert _rel(acc - 1.0, ref_dw) < 6e-5
    # bitwise reproducible (fixed issue order of the MMAs)
    y2, _ = m.forward_with_tap(xr)
    assert torch.equal(y, y2)
    # zero padding: an all-ones kernel on a constant image counts the taps inside the image (exact in TF32)
    with torch.no_grad():
        m.weight.fill_(1.0)
    ones = m(torch.ones(2, c, h, h, device=dev))
    assert ones[0, 0, 0, 0].item() == 4 * c and ones[1, 0, 1, 1].item() == 9 * c and ones[0, -1, -1, 0].item() == 4 * c
    assert ones[1, 0, 0, 1].item() == 6 * c and ones[1, c // 2, h - 1, h - 1].item() == 4 * c


def test_tcgen05_mode_falls_back_outside_its_shapes(monkeypatch):
    """tc3 mode: C = 16 layers run the FFMA kernel, C in {32, 64} on maps the tcgen05 kernel does not cover (or an odd
    batch of 8x8 maps) run the library convolution -- never a wrong result."""
    monkeypatch.setattr(conv, "MODE", "tc3")
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)       # the library fallback in strict fp32 too
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for n, c, h in [(4, 16, 32), (5, 64, 8), (3, 32, 8), (2, 64, 16)]:
        m = conv.Conv3x3(c, c, 1).to(dev)
        x = torch.randn(n, c, h, h, device=dev)
        ref = F.conv2d(x.double(), m.weight.double(), padding=1)
        assert _rel(m(x), ref) < 6e-5, (n, c, h)


@pytest.mark.parametrize("n,c,h,groups", [(128, 32, 16, 1), (128, 64, 8, 1), (256, 32, 16, 2), (256, 64, 8, 2), (6, 32, 16, 2), (4, 64, 8, 1)])
def test_tcgen05_conv_with_folded_batchnorm(n, c, h, groups):
    """BatchNorm folded into the tcgen05 convolution (resnet_s.py:70-72): statistics of the output in the epilogue
    (= the stand-alone dual-BN kernel on the same tensor), the producer's BatchNorm + ReLU applied on load (= convolving the
    materialised activation), and the BatchNorm + ReLU backward with the mask recomputed from x."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(n + c + h + groups)
    x = torch.randn(n, c, h, h, generator=g).to(dev)
    dy = torch.randn(n, c, h, h, generator=g).to(dev)
    m1, m2 = conv.Conv3x3(c, c, 1).to(dev), conv.Conv3x3(c, c, 1).to(dev)
    descs = torch.tensor([m1.desc_row(), m2.desc_row()], dtype=torch.int64, device=dev)
    ops.conv3x3_pack(descs, c, "umma")
    wf1, wf2 = m1._packed[0], m2._packed[0]
    w, b = (torch.rand(c, generator=g) + 0.5).to(dev), torch.randn(c, generator=g).to(dev)
    rm0, rv0 = torch.randn(c, generator=g).to(dev), (torch.rand(c, generator=g) + 0.5).to(dev)
    # un-fused: conv -> dual BN (+ReLU) kernel -> conv
    c1 = ops.conv3x3(x, wf1, math="umma")
    rm_a, rv_a = rm0.clone(), rv0.clone()
    hmat, sm, si = ops.bn_fwd(c1, None, w, b, rm_a, rv_a, ops.bn_workspace(groups, c, dev), groups=groups, relu=True, replay=2)
    c2 = ops.conv3x3(hmat, wf2, math="umma")
    dx_ref, _, dw_ref, db_ref = ops.bn_bwd(dy, c1, hmat, w, sm, si, ops.bn_workspace(groups, c, dev), groups=groups, relu=True)
    # fused: the producer writes per-CTA statistics, the consumer folds them and normalises while loading
    rm_b, rv_b = rm0.clone(), rv0.clone()
    ws = ops.conv3x3_umma_bn_workspace(n, c, dev)
    for _ in range(2):                                    # twice: the workspace is reusable as is
        rm_b.copy_(rm0); rv_b.copy_(rv0)
        c1f, _, _, _ = ops.conv3x3_umma_bn(x, wf1, stats_out=ws, groups=groups)
        c2f, smf, sif, tab = ops.conv3x3_umma_bn(c1f, wf2, stats_in=ws, bn=(w, b, rm_b, rv_b), groups=groups, replay=2)
    assert torch.equal(c1f, c1)                           # same MMA sequence
    torch.testing.assert_close(smf, sm, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(sif, si, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rm_b, rm_a, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv_b, rv_a, rtol=1e-5, atol=1e-6)
    assert _rel(c2f, c2.double()) < 1e-5
    dxf, dwf, dbf = ops.bn_bwd_xmask(dy, c1f, tab, w, smf, sif, groups=groups)
    assert _rel(dxf, dx_ref.double()) < 1e-4
    torch.testing.assert_close(dwf, dw_ref, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dbf, db_ref, rtol=1e-4, atol=1e-3)
    # deterministic: fixed-order fold
    c2g, smg, sig, tabg = ops.conv3x3_umma_bn(c1f, wf2, stats_in=ws, bn=(w, b, rm_b.clone(), rv_b.clone()), groups=groups)
    assert torch.equal(smg, smf) and torch.equal(tabg, tab) and torch.equal(c2g, c2f)


@pytest.mark.parametrize("n,c,h,groups", [(128, 32, 16, 1), (128, 64, 8, 1), (8, 32, 16, 2)])
def test_fused_basic_block_equals_the_unfused_block(n, c, h, groups, monkeypatch):
    """AFAN_FUSE_BN1: conv1 -> [bn1 + relu folded into conv2] -> conv2 -> bn2 (+shortcut, relu) with frozen parameters (a
    PGD ascent pass) against the same block with its four separate launches: output, input gradient, running statistics."""
    rs = pkg.resnet_s
    dev = torch.device("cuda:0")
    monkeypatch.setattr(conv, "MODE", "tc3")
    torch.manual_seed(n + c)
    blk = rs.BasicBlock(c, c, 1).to(dev).train()
    with torch.no_grad():
        for bn in (blk.bn1, blk.bn2):
            bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    for p in blk.parameters():
        p.requires_grad_(False)
    x = torch.randn(n, c, h, h, device=dev)
    dy = torch.randn(n, c, h, h, device=dev)
    state0 = {k: v.clone() for k, v in blk.state_dict().items()}
    outs = {}
    for fuse in (False, True):
        monkeypatch.setattr(rs, "FUSE_BN1", fuse)
        blk.load_state_dict(state0)
        blk.bn1._pending_batches = blk.bn2._pending_batches = 0
        xr = x.clone().requires_grad_(True)
        assert blk._fusable(xr, groups) == fuse
        y = blk(xr, groups=groups, replay=1)
        (dx,) = torch.autograd.grad(y, xr, dy)
        outs[fuse] = (y.detach(), dx, {k: v.clone() for k, v in blk.state_dict().items()})
    (y0, dx0, s0), (y1, dx1, s1) = outs[False], outs[True]
    assert _rel(y1, y0.double()) < 2e-5 and _rel(dx1, dx0.double()) < 2e-4
    for k in s0:
        if s0[k].dtype.is_floating_point:
            torch.testing.assert_close(s1[k], s0[k], rtol=1e-5, atol=1e-6, msg=lambda m, k=k: f"{k}: {m}")
        else:
            assert torch.equal(s1[k], s0[k]), k


@pytest.mark.parametrize("n_loc,c,h,groups", [(16, 32, 16, 1), (16, 64, 8, 1), (8, 32, 16, 2)])
def test_folded_batchnorm_exchange_loopback_two_virtual_ranks(n_loc, c, h, groups):
    """Multi-GPU form of the folded BatchNorm, exercised on ONE GPU like test_fused_p2p_exchange_loopback…: two virtual
    ranks on two streams with mailboxes in local memory.  Each rank's consumer convolution folds its local producer
    statistics, exchanges them inside its prologue and normalises with the statistics of the GLOBAL batch; the oracle is
    the single-process un-fused sequence (conv -> dual-BN kernel -> conv, and its backward) over the global batch."""
    dev = torch.device("cuda:0")
    world = 2
    g = torch.Generator(device="cpu").manual_seed(n_loc + c)
    n_glob = n_loc * world                                   # per statistic group
    x = torch.randn(groups * n_glob, c, h, h, generator=g).to(dev)
    dh = torch.randn(groups * n_glob, c, h, h, generator=g).to(dev)
    shard = lambda t, r: torch.cat([t[gi * n_glob + r * n_loc: gi * n_glob + (r + 1) * n_loc] for gi in range(groups)]).contiguous()
    m1, m2 = conv.Conv3x3(c, c, 1).to(dev), conv.Conv3x3(c, c, 1).to(dev)
    ops.conv3x3_pack(torch.tensor([m1.desc_row(), m2.desc_row()], dtype=torch.int64, device=dev), c, "umma")
    wf1, wf2 = m1._packed[0], m2._packed[0]
    w, b = (torch.rand(c, generator=g) + 0.5).to(dev), torch.randn(c, generator=g).to(dev)
    # oracle: global batch, one process
    c1g = ops.conv3x3(x, wf1, math="umma")
    rm_g, rv_g = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    hg, smg, sig = ops.bn_fwd(c1g, None, w, b, rm_g, rv_g, ops.bn_workspace(groups, c, dev), groups=groups, relu=True)
    c2g = ops.conv3x3(hg, wf2, math="umma")
    dc1g, _, dwg, dbg = ops.bn_bwd(dh, c1g, hg, w, smg, sig, ops.bn_workspace(groups, c, dev), groups=groups, relu=True)
    # two virtual ranks; every exchanging kernel is launched for both ranks before anything host-synchronous happens
    boxes = pkg.p2p.PeerMailbox.loopback(world, dev, cmax=c)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    xs, dhs = [shard(x, r) for r in range(world)], [shard(dh, r) for r in range(world)]
    wss = [ops.conv3x3_umma_bn_workspace(groups * n_loc, c, dev) for _ in range(world)]
    rms, rvs = [torch.zeros(c, device=dev) for _ in range(world)], [torch.ones(c, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    c1, fw, bw = [None] * world, [None] * world, [None] * world
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            c1[r] = ops.conv3x3_umma_bn(xs[r], wf1, stats_out=wss[r], groups=groups)[0]
    torch.cuda.synchronize()
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            fw[r] = ops.conv3x3_umma_bn(c1[r], wf2, stats_in=wss[r], bn=(w, b, rms[r], rvs[r]), groups=groups, mailbox=boxes[r])
    torch.cuda.synchronize()
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            bw[r] = ops.bn_bwd_xmask(dhs[r], c1[r], fw[r][3], w, fw[r][1], fw[r][2], groups=groups, mailbox=boxes[r])
    torch.cuda.synchronize()
    for bx in boxes:
        bx.check()
    assert int(boxes[0].state[0]) == 2 and int(boxes[1].state[0]) == 2                 # two exchanges counted on each rank
    assert torch.equal(fw[0][1], fw[1][1]) and torch.equal(fw[0][2], fw[1][2]) and torch.equal(fw[0][3], fw[1][3])   # bit-identical statistics
    dw_sum, db_sum = 0, 0
    for r in range(world):
        c2, sm, si, tab = fw[r]
        torch.testing.assert_close(sm, smg, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(si, sig, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(rvs[r], rv_g, rtol=1e-5, atol=1e-6)
        assert _rel(c2, shard(c2g, r).double()) < 1e-5
        assert _rel(bw[r][0], shard(dc1g, r).double()) < 1e-4
        dw_sum, db_sum = dw_sum + bw[r][1], db_sum + bw[r][2]
    torch.testing.assert_close(dw_sum, dwg, rtol=1e-4, atol=1e-3)                      # local dweight / dbias sum to the global ones
    torch.testing.assert_close(db_sum, dbg, rtol=1e-4, atol=1e-3)
    for bx in boxes:
        bx.close()
