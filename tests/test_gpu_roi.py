"""-m gpu: ROIAlign forward / backward vs the oracle restatement of the reference kernel."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import PKG, dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,r,out,scale,ratio", [((2, 8, 38, 63), 16, (14, 14), 1 / 16, 0),     # Faster R-CNN pooler (roi/pooler.py:35)
                                                     ((1, 3, 10, 12), 5, (7, 7), 1 / 16, 2), ((2, 4, 5, 5), 3, (2, 3), 1.0, 0),
                                                     # round-2 paths: ragged channel chunks (40 = 32 + 8), an image without any
                                                     # ROI (index 2 of 3 never drawn below -> zero gradient plane), many ROIs
                                                     ((3, 40, 38, 63), 48, (14, 14), 1 / 16, 0),
                                                     # > 16 sample points per bin axis and planes too large for shared memory:
                                                     # on-the-fly taps / one-thread-per-element kernels
                                                     ((1, 5, 260, 250), 4, (7, 7), 1.0, 0)])
def test_roi_align_fwd_bwd_vs_oracle(shape, r, out, scale, ratio):
    g = torch.Generator().manual_seed(r)
    n, c, h, w = shape
    feat = torch.randn(shape, generator=g)
    img_w, img_h = w / scale, h / scale
    x1, y1 = torch.rand(r, generator=g) * img_w * 0.8, torch.rand(r, generator=g) * img_h * 0.8
    batch = torch.randint(0, max(n - 1, 1) if n == 3 else n, (r,), generator=g).float()      # n == 3: image 2 gets no ROI
    rois = torch.stack([batch, x1, y1, x1 + torch.rand(r, generator=g) * img_w * 0.6,
                        y1 + torch.rand(r, generator=g) * img_h * 0.6], 1)
    if h >= 200:
        rois[1, 1:] = torch.tensor([2.0, 3.0, img_w - 4.0, img_h - 5.0])   # nearly the whole map: ceil(250 / 7) = 36 points per bin axis
    rois[0, 1:] = torch.tensor([-30.0, -20.0, 10.0, 15.0])              # partly outside the map
    rois[-1, 3:] = rois[-1, 1:3]                                        # degenerate -> forced to 1x1
    dy = torch.randn(r, c, *out, generator=g)
    ft = feat.to(dev()).requires_grad_(True)
    layer = PKG.detection.ROIAlign(out, scale, ratio)
    y = layer(ft, rois.to(dev()))
    y.backward(dy.to(dev()))
    y_ref, d_ref = orc.roi_align(feat.numpy(), rois.numpy(), out, scale, ratio, dout=dy.numpy())
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_ref, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ft.grad.cpu().numpy(), d_ref, rtol=1e-4, atol=5e-5)   # scatter: summation order differs
    if n == 3:
        assert float(ft.grad[2].abs().max()) == 0.0                     # plane-resident backward writes zeros itself (no memset)


def test_roi_align_empty_rois():
    ft = torch.randn(1, 2, 4, 4, device=dev())
    y = PKG.detection.roi_align(ft, torch.zeros(0, 5, device=dev()), (2, 2), 1.0, 0)
    assert y.shape == (0, 2, 2, 2)
