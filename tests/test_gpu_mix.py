"""-m gpu: mix_feature kernel vs reference goldens (Seg + Det flavours) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import GOLDEN, PKG, dev, feature_like

pytestmark = pytest.mark.gpu


def test_mix_feature_reference_goldens():
    z = np.load(os.path.join(GOLDEN, "helpers.npz"))
    for i in range(int(z["n_mix"])):
        cl, ad = torch.from_numpy(z[f"mix{i}_clean"]).to(dev()), torch.from_numpy(z[f"mix{i}_adv"]).to(dev())
        for mod, key in ((PKG.segmentation, "seg"), (PKG.detection, "det")):
            got = mod.mix_feature(cl, ad).cpu().numpy()
            np.testing.assert_allclose(got, z[f"mix{i}_{key}"], rtol=2e-5, atol=2e-6, err_msg=f"mix{i} {key}")
        for n in (3, 5):
            pts = PKG.segmentation.get_sample_points(cl, ad, n)
            for j in range(n):
                np.testing.assert_allclose(pts[j].cpu().numpy(), z[f"mix{i}_pts{n}"][j], rtol=0, atol=1.2e-7)


@pytest.mark.parametrize("shape", [(2, 1024, 33, 33), (4, 256, 33, 33), (2, 2048, 8, 8), (3, 304, 12, 20),
                                   (1, 2, 3, 3), (2, 17, 1, 5), (1, 1, 2, 2), (2, 40, 28, 28)])
def test_mix_feature_vs_oracle(shape):
    g = torch.Generator().manual_seed(sum(shape))
    cl = feature_like(shape, g)
    ad = cl + (2 / 255) * torch.sign(torch.randn(shape, generator=g))
    got = PKG.ops.mix_feature(cl.to(dev()), ad.to(dev())).cpu().numpy()
    ref = orc.mix_feature(cl.numpy(), ad.numpy())
    assert np.array_equal(np.isnan(got), np.isnan(ref))          # C == 1 -> NaN like torch.var
    m = ~np.isnan(ref)
    np.testing.assert_allclose(got[m], ref[m], rtol=2e-5, atol=2e-6)


def test_mix_feature_outlier_channel_is_stable():
    """One huge channel value per pixel: a sum/sum-of-squares formulation would cancel catastrophically."""
    g = torch.Generator().manual_seed(0)
    cl = 0.01 * torch.randn(2, 64, 6, 6, generator=g)
    cl[:, 0] = 100.0
    ad = cl + 0.005 * torch.randn(cl.shape, generator=g)
    got = PKG.ops.mix_feature(cl.to(dev()), ad.to(dev())).cpu().numpy()
    np.testing.assert_allclose(got, orc.mix_feature(cl.numpy(), ad.numpy()), rtol=1e-4, atol=1e-4)


def test_mix_identity_property_full_size():
    """mix_feature(x, x) == x up to rounding, at a DeepLab-sized tensor (size-independent property)."""
    g = torch.Generator().manual_seed(1)
    x = feature_like((4, 2048, 33, 33), g).to(dev())
    out = PKG.ops.mix_feature(x, x)
    torch.testing.assert_close(out, x, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("shape", [(2, 256, 33, 33), (2, 64, 16, 16), (1, 5, 3, 7)])
@pytest.mark.parametrize("number,mix", [(3, (True, True)), (3, (False, True)), (5, (True, False, True, False)),
                                        (5, (False, False, False, False)), (5, (True, True, True, True))])
def test_fused_sat_points_vs_oracle(shape, number, mix):
    """get_sample_points + per-point mix_feature (main_aug_final.py:206-210, train_aug_final.py:117-126) in one launch."""
    g = torch.Generator().manual_seed(number + sum(shape))
    cl = feature_like(shape, g)
    ad = cl + (2 / 255) * torch.sign(torch.randn(shape, generator=g))
    cl_d, ad_d = cl.to(dev()), ad.to(dev())
    got = PKG.segmentation.sat_sample_points(cl_d, ad_d, number, mix)
    ref_pts = orc.get_sample_points(cl.numpy(), ad.numpy(), number)
    assert len(got) == number and got[0] is cl_d
    for i in range(1, number):
        exp = orc.mix_feature(cl.numpy(), ref_pts[i]) if mix[i - 1] else ref_pts[i]
        if mix[i - 1]:
            np.testing.assert_allclose(got[i].cpu().numpy(), exp, rtol=2e-5, atol=2e-6, err_msg=f"point {i}")
        else:
            np.testing.assert_allclose(got[i].cpu().numpy(), exp, rtol=0, atol=1.2e-7, err_msg=f"point {i}")
    if not mix[-1]:
        assert got[-1] is ad_d                       # the reference returns pointy itself
    # the un-fused interface gives the same points
    plain = PKG.segmentation.get_sample_points(cl_d, ad_d, number)
    for i in range(1, number - 1):
        if not mix[i - 1]:
            np.testing.assert_allclose(got[i].cpu().numpy(), plain[i].cpu().numpy(), rtol=0, atol=1.2e-7)
