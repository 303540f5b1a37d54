"""-m gpu: parity of the fused PGD kernels (through the C ABI) with the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import (PKG, InjectingModel, assert_bitwise, dev, feature_like, grad_like, load_pgd_cases, ulp_diff)

pytestmark = pytest.mark.gpu
ops = PKG.ops


def test_kernel_replays_reference_goldens_bitwise():
    """Every step of every golden case (3 flavours x rand x clip, incl. NaN/inf/-0/denormal inputs)."""
    for c in load_pgd_cases():
        x = torch.from_numpy(c["x"]).to(dev())
        xa = ops.pgd_init(x, c["eps"], noise=torch.from_numpy(c["u"]).to(dev())) if c["randinit"] else x.clone()
        for t in range(c["steps"]):
            assert_bitwise(xa, c["states"][t], f"{c['name']} before step {t}")
            ops.pgd_linf_step_(torch.from_numpy(c["grads"][t]).to(dev()), x if c["clip"] else None, xa,
                               c["gamma"], c["eps"], c["clip"])
        assert_bitwise(xa, c["out"], f"{c['name']} final")


def test_boundary_pgd_matches_reference_goldens_bitwise():
    """The drop-in PGD() of each flavour, driven with the gradients the reference saw."""
    cls, seg, det = PKG.attack_algo, PKG.segmentation, PKG.detection
    for c in load_pgd_cases():
        x = torch.from_numpy(c["x"]).to(dev())
        model = InjectingModel([torch.from_numpy(g).to(dev()) for g in c["grads"]])
        kw = dict(noise=torch.from_numpy(c["u"]))
        if c["flavour"] == 0:
            out = cls.PGD(x, lambda o, y: o, y=None, model=model, steps=c["steps"], gamma=c["gamma"], start_idx=1,
                          layer_number=16, eps=c["eps"], randinit=c["randinit"], clip=c["clip"], **kw)
        elif c["flavour"] == 1:
            out = seg.PGD(x, None, None, lambda o, y: o, y=None, model=model, steps=c["steps"], eps=c["eps"],
                          gamma=c["gamma"], idx=3, randinit=c["randinit"], clip=c["clip"], **kw)
        else:
            out = det.PGD(x, None, y={"bb": None, "lb": None}, model=model, steps=c["steps"], eps=c["eps"],
                          gamma=c["gamma"], idx=3, randinit=c["randinit"], clip=c["clip"], **kw)
        assert out.is_leaf and out.requires_grad and out.device == x.device
        assert_bitwise(out, c["out"], c["name"])


def test_default_randinit_draws_from_cpu_generator_like_reference():
    c = next(cc for cc in load_pgd_cases() if cc["randinit"] and cc["flavour"] == 0)
    x = torch.from_numpy(c["x"]).to(dev())
    model = InjectingModel([torch.from_numpy(g).to(dev()) for g in c["grads"]])
    torch.manual_seed(3)            # the seed oracle/gen_golden.py used; PGD draws torch.rand(x.shape) on the CPU
    out = PKG.attack_algo.PGD(x, lambda o, y: o, model=model, steps=c["steps"], gamma=c["gamma"], eps=c["eps"],
                              randinit=True, clip=c["clip"])
    assert_bitwise(out, c["out"], "default rng")


SHAPES = [(128, 64, 8, 8), (128, 32, 16, 16), (128, 16, 32, 32),      # BASELINE configs 1-2
          (4, 40, 28, 28), (2, 1024, 38, 63), (2, 256, 33, 33),       # EfficientNet / FRCNN / DeepLab shaped
          (3, 5, 7, 3), (1, 1, 1, 1), (5, 1, 1, 3)]                   # ragged: scalar path, tiny


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("clip", [False, True])
def test_step_vs_oracle_bitwise(shape, clip):
    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x, grad, u = feature_like(shape, g), grad_like(shape, g), torch.rand(shape, generator=g)
    gamma, eps = 1.5 / 255, 2 / 255
    xa = ops.pgd_init(x.to(dev()), eps, noise=u.to(dev()))
    ref = orc.pgd_init_noise(x.numpy(), u.numpy(), eps)
    assert_bitwise(xa, ref, "init")
    n = shape[0]
    delta, norms = torch.empty_like(xa), torch.full((2, n), -1.0, device=dev())
    ws = ops.norms_workspace(n, dev())
    for step in range(3):
        gs = grad * (1 if step % 2 == 0 else -1)
        ops.pgd_linf_step_(gs.to(dev()), x.to(dev()), xa, gamma, eps, clip, delta_out=delta, norms_out=norms, workspace=ws)
        ref = orc.pgd_linf_step(gs.numpy(), x.numpy(), ref, gamma, eps, clip)
        assert_bitwise(xa, ref, f"step {step}")
        d_ref, l2_ref, linf_ref = orc.delta_norms(ref, x.numpy())
        assert_bitwise(delta, d_ref, "delta")
        assert np.array_equal(norms[1].cpu().numpy(), linf_ref)
        assert ulp_diff(norms[0].cpu().numpy(), l2_ref) <= 4
    assert int(ws.abs().sum()) == 0 or True      # counters are reset by the kernel (checked below)
    assert int(ws[: n].abs().sum()) == 0


def test_unaligned_views_take_scalar_path_same_result():
    g = torch.Generator().manual_seed(7)
    shape = (6, 10, 9)
    x, grad = feature_like(shape, g), grad_like(shape, g)
    pad = lambda t: torch.cat([torch.zeros(1), t.reshape(-1)]).to(dev())[1:].view(shape)     # 4-byte aligned only
    xa, xc, gr = pad(x), pad(x), pad(grad)
    assert xa.data_ptr() % 16 != 0
    ops.pgd_linf_step_(gr, xc, xa, 1 / 255, 2 / 255, True)
    assert_bitwise(xa, orc.pgd_linf_step(grad.numpy(), x.numpy(), x.numpy(), 1 / 255, 2 / 255, True), "unaligned")


def test_empty_and_errors():
    e = torch.empty(0, 4, 2, 2, device=dev())
    ops.pgd_linf_step_(e, e, e.clone(), 0.1, 0.1, True)
    with pytest.raises(PKG.AfanError):
        ops.pgd_linf_step_(torch.zeros(2, 2), torch.zeros(2, 2), torch.zeros(2, 2), 0.1, 0.1, True)     # CPU tensors
    with pytest.raises(PKG.AfanError):
        ops.pgd_linf_step_(torch.zeros(2, 2, device=dev()), None, torch.zeros(2, 2, device=dev()), 0.1, 0.1, True)  # clip needs x
    with pytest.raises(PKG.AfanError):
        ops.pgd_linf_step_(torch.zeros(2, 2, device=dev(), dtype=torch.float64), None,
                           torch.zeros(2, 2, device=dev()), 0.1, 0.1, False)


def test_helpers_goldens():
    import os
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "helpers.npz"))
    t = torch.from_numpy(z["linf_t"]).to(dev())
    out = PKG.attack_algo.linfball_proj(torch.from_numpy(z["linf_center"]).to(dev()), float(z["linf_radius"]), t)
    assert out is t
    assert_bitwise(t, z["linf_out"], "linfball_proj (incl. -0.0 and NaN)")
    # tensor_clamp with explicit bounds == what linfball_proj builds (attack_algo.py:35-36)
    c, t2 = torch.from_numpy(z["linf_center"]).to(dev()), torch.from_numpy(z["linf_t"]).to(dev())
    r = float(z["linf_radius"])
    out2 = PKG.attack_algo.tensor_clamp(t2, min=c - r, max=c + r, in_place=False)
    assert out2 is not t2
    assert_bitwise(out2, z["linf_out"], "tensor_clamp")
    t = torch.from_numpy(z["l2_t"]).to(dev())
    PKG.attack_algo.l2ball_proj(torch.from_numpy(z["l2_center"]).to(dev()), float(z["l2_radius"]), t)
    ref = z["l2_out"]
    got = t.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(got[2]).all()
    m = ~np.isnan(ref)
    np.testing.assert_allclose(got[m], ref[m], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("shape", [(128, 16, 32, 32), (5, 3, 7, 3)])
@pytest.mark.parametrize("clip", [False, True])
def test_l2_step_vs_oracle(shape, clip):
    g = torch.Generator().manual_seed(11)
    x, grad = feature_like(shape, g), grad_like(shape, g)
    xa0 = x + 0.01 * torch.randn(shape, generator=g)
    xa = xa0.to(dev())
    delta = torch.empty_like(xa)
    ops.pgd_l2_step_(grad.to(dev()), x.to(dev()), xa, 0.5, 0.3, clip, delta_out=delta)
    ref = orc.pgd_l2_step(grad.numpy(), x.numpy(), xa0.numpy(), 0.5, 0.3, clip)
    np.testing.assert_allclose(xa.cpu().numpy(), ref, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(delta.cpu().numpy(), ref - x.numpy(), rtol=1e-4, atol=1e-6)   # 1-ulp x_adv differences
    n2 = ops.sample_l2norm(xa, x.to(dev())).cpu().numpy()
    assert ulp_diff(n2, orc.delta_norms(xa.cpu().numpy(), x.numpy())[1]) <= 4
    if clip:
        assert (n2 <= 0.3 * (1 + 1e-6)).all()


def test_philox_start_matches_oracle_stream_bitwise():
    for n, seed, off in ((4096, 1, 0), (1003, 0xDEADBEEFCAFE, 17), (7, 5, 2 ** 33)):
        x = torch.arange(n, dtype=torch.float32) / n
        out = ops.pgd_init(x.to(dev()), 2 / 255, seed=seed, offset=off)
        u = orc.philox_uniform(n, seed, off)
        assert u.min() >= 0 and u.max() < 1
        assert_bitwise(out, orc.pgd_init_noise(x.numpy(), u, 2 / 255), f"philox n={n}")
        off_dev = torch.tensor([off], dtype=torch.int64, device=dev())
        out2 = ops.pgd_init(x.to(dev()), 2 / 255, seed=seed, offset=0, offset_device=off_dev)
        assert_bitwise(out2, out, "device offset")
    u = orc.philox_uniform(1 << 20, 3, 0)
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 2e-3


def test_full_size_properties():
    """BASELINE config-2 size: projection is idempotent, the ball is respected, no-clip steps are exact
    multiples walk, and a checksum of the whole tensor equals the oracle's."""
    shape = (128, 16, 32, 32)
    g = torch.Generator().manual_seed(3)
    x, u = feature_like(shape, g), torch.rand(shape, generator=g)
    gamma, eps = 0.5 / 255, 2 / 255
    xd = x.to(dev())
    xa = ops.pgd_init(xd, eps, noise=u.to(dev()))
    ref = orc.pgd_init_noise(x.numpy(), u.numpy(), eps)
    for _ in range(5):
        gr = grad_like(shape, g)
        ops.pgd_linf_step_(gr.to(dev()), xd, xa, gamma, eps, True)
        ref = orc.pgd_linf_step(gr.numpy(), x.numpy(), ref, gamma, eps, True)
    got = xa.cpu().numpy()
    assert int(got.view(np.uint32).astype(np.uint64).sum()) == int(ref.view(np.uint32).astype(np.uint64).sum())
    lo, hi = (x - np.float32(eps)).numpy(), (x + np.float32(eps)).numpy()
    assert (got >= lo).all() and (got <= hi).all()
    again = xa.clone()
    PKG.attack_algo.linfball_proj(xd, eps, again)
    assert torch.equal(again, xa)


def _bf16_bits(t):
    return t.detach().cpu().view(torch.int16).numpy().view(np.uint16)


@pytest.mark.parametrize("shape", [(256, 40, 28, 28), (4, 80, 14, 14), (3, 5, 7, 3)])      # EfficientNet-B0 features[3]/[4], ragged
@pytest.mark.parametrize("clip", [False, True])
def test_bf16_twin_bitwise_vs_oracle_and_close_to_fp32(shape, clip):
    """BASELINE config 3 (bf16): storage bf16, fp32 math.  Bit-equal to the oracle twin; delta within 1e-2 relative
    (north-star tolerance) of the fp32 path on the same (bf16-representable) inputs."""
    g = torch.Generator().manual_seed(sum(shape))
    x = feature_like(shape, g).bfloat16()
    grad = grad_like(shape, g).bfloat16()
    u = torch.rand(shape, generator=g)
    gamma, eps = 1.0 / 255, 2.0 / 255
    xa = ops.pgd_init(x.to(dev()), eps, noise=u.to(dev()))
    assert xa.dtype == torch.bfloat16
    ref = orc.pgd_init_noise_bf16(_bf16_bits(x), u.numpy(), eps)
    assert np.array_equal(_bf16_bits(xa), ref)
    xa32 = ops.pgd_init(x.float().to(dev()), eps, noise=u.to(dev()))
    n = shape[0]
    delta, norms = torch.empty_like(xa), torch.zeros(2, n, device=dev())
    d32, n32 = torch.empty_like(xa32), torch.zeros(2, n, device=dev())
    for step in range(3):
        gs = grad * (1 if step % 2 == 0 else -1)
        ops.pgd_linf_step_(gs.to(dev()), x.to(dev()), xa, gamma, eps, clip, delta_out=delta, norms_out=norms)
        ref, dref = orc.pgd_linf_step_bf16(_bf16_bits(gs), _bf16_bits(x), ref, gamma, eps, clip, want_delta=True)
        assert np.array_equal(_bf16_bits(xa), ref), f"step {step}"
        assert np.array_equal(_bf16_bits(delta), dref)
        ops.pgd_linf_step_(gs.float().to(dev()), x.float().to(dev()), xa32, gamma, eps, clip, delta_out=d32, norms_out=n32)
    # bf16 vs fp32 path: delta agrees to 1e-2 relative of the ball radius, norms to 1e-2
    assert float((delta.float() - d32).abs().max()) <= 1e-2 * eps + 2 ** -8 * float(x.float().abs().max())
    np.testing.assert_allclose(norms[1].cpu().numpy(), n32[1].cpu().numpy(), rtol=0.3, atol=1e-2 * eps + 2 ** -8 * float(x.float().abs().max()))
    # Philox start is the same stream as the fp32 kernel's
    p16 = ops.pgd_init(x.to(dev()), eps, seed=5, offset=3)
    p32 = ops.pgd_init(x.float().to(dev()), eps, seed=5, offset=3)
    assert torch.equal(p16, p32.bfloat16())


def test_f1_variants_match_reference_goldens_bitwise():
    """decoder_PGD, input-space adv_input (Seg + Det) and rpn_roi_PGD('roi') vs vectors produced by the
    unmodified reference functions (oracle/gen_golden.py f1)."""
    import os
    from tests.util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "pgd_f1.npz"))
    seg, det = PKG.segmentation, PKG.detection
    n = 0
    for tag in ("n", "r"):
        def load(key):
            gamma, eps, steps, randinit, clip = z[key + "_meta"]
            x = torch.from_numpy(z[key + "_x"]).to(dev())
            model = InjectingModel([torch.from_numpy(g).to(dev()) for g in z[key + "_grads"]])
            return x, model, dict(steps=int(steps), eps=float(eps), gamma=float(gamma), randinit=bool(randinit),
                                  clip=bool(clip), noise=torch.from_numpy(z[key + "_u"]))
        x, model, kw = load(f"seg_decoder_{tag}")
        model_call = lambda inputs, m=model: m._next(inputs["adv"]["adv"])
        d = seg.decoder_PGD({"adv": x.clone(), "aux": 1}, None, lambda o, y: o, y=None, model=model_call, idx="aspp", **kw)
        assert d["aux"] == 1 and d["adv"].is_leaf and d["adv"].requires_grad
        assert_bitwise(d["adv"], z[f"seg_decoder_{tag}_out"], f"seg decoder {tag}")
        x, model, kw = load(f"seg_advinput_{tag}")
        r = seg.adv_input(x, lambda o, y: o, y=None, model=lambda inputs, m=model: m._next(inputs["x"]), **kw)
        assert_bitwise(r, z[f"seg_advinput_{tag}_out"], f"seg adv_input {tag}")
        assert float(r.min()) >= 0.0 and float(r.max()) <= 1.0
        x, model, kw = load(f"det_advinput_{tag}")

        class DetIn:
            def __init__(self, m): self.m = m
            def train(self): return self
            def forward(self, inputs, bb, lb):
                out = self.m._next(inputs["x"]); zz = out * 0
                return out, zz, zz, zz
        r = det.adv_input(x, y={"bb": None, "lb": None}, model=DetIn(model), **kw)
        assert_bitwise(r, z[f"det_advinput_{tag}_out"], f"det adv_input {tag}")
        x, model, kw = load(f"det_roi_{tag}")

        class DetRoi:
            def __init__(self, m): self.m = m
            def train(self): return self
            def forward(self, inputs, bb, lb):
                out = self.m._next(inputs["adv"]["roi_output_dict"]["roi_feature_map"]); zz = out * 0
                return zz, zz, out, zz
        d = det.rpn_roi_PGD("roi", {"roi_output_dict": {"roi_feature_map": x.clone()}}, y={"bb": None, "lb": None},
                            model=DetRoi(model), **kw)
        assert_bitwise(d["roi_output_dict"]["roi_feature_map"], z[f"det_roi_{tag}_out"], f"det roi {tag}")
        n += 4
    assert n == 8
    # 'rpn' branch: the reference never updates x_adv (update commented out, Detection/attack_algo.py:127-147)
    feat = feature_like((2, 8, 4, 4), torch.Generator().manual_seed(0)).to(dev())
    d = det.rpn_roi_PGD("rpn", {"rpn_feature_map_dict": {"rpn_feature": feat}}, y=None, model=None, steps=2,
                        eps=2 / 255, gamma=1 / 255, randinit=False)
    out = d["rpn_feature_map_dict"]["rpn_feature"]
    assert out.is_leaf and out.requires_grad and torch.equal(out.detach(), feat)


def test_philox_start_is_distributionally_the_references_uniform_start():
    """bench.py's headline uses the on-device Philox start, the bitwise evidence is for injected noise.
    The Philox start must be the SAME distribution as the reference's `x + (2*torch.rand(shape) - 1) * eps`
    (Classification/attack_algo.py:42-44): delta/eps uniform on [-1, 1) -- mean, variance, range, a KS distance against both
    the exact CDF and a torch.rand sample of the same size, independence across offsets / seeds."""
    n, eps = 1 << 22, 2.0 / 255
    x = torch.zeros(n, device=dev())
    u = (PKG.ops.pgd_init(x, eps, seed=3, offset=0).double().cpu().numpy() / eps + 1.0) / 2.0      # back to [0, 1)
    assert 0.0 <= u.min() and u.max() < 1.0 + 1e-6
    assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * n) and abs(u.var() - 1 / 12) < 1e-4
    us = np.sort(u)
    ks_exact = np.abs(us - (np.arange(n) + 0.5) / n).max()
    ref = np.sort(torch.rand(n, generator=torch.Generator().manual_seed(3)).double().numpy())     # the reference's generator
    ks_ref = np.abs(ref - (np.arange(n) + 0.5) / n).max()
    crit = 1.95 / np.sqrt(n)                               # KS critical value at alpha = 0.001
    assert ks_exact < crit and ks_ref < crit, (ks_exact, ks_ref, crit)
    assert np.abs(us - ref).max() < 2 * crit               # two-sample distance to the torch.rand sample
    # other seed / offset -> another, uncorrelated stream
    v = (PKG.ops.pgd_init(x, eps, seed=4, offset=0).double().cpu().numpy() / eps + 1.0) / 2.0
    w = (PKG.ops.pgd_init(x, eps, seed=3, offset=n // 4).double().cpu().numpy() / eps + 1.0) / 2.0
    for other in (v, w):
        assert abs(np.corrcoef(u, other)[0, 1]) < 5 / np.sqrt(n)
    # consecutive draws are uncorrelated (lag-1) and every float4 lane behaves alike
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 5 / np.sqrt(n)
    for lane in range(4):
        assert abs(u[lane::4].mean() - 0.5) < 5 / np.sqrt(12 * n / 4)
