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
