"""-m gpu: the small launches of the Classification tail -- option-A shortcut (resnet_s.py:60-63) and the classifier's weight /
bias gradient (resnet_s.py:93-95) -- against the library sequences they replace."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import PKG, dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,pad", [((4, 16, 32, 32), 8), ((3, 32, 16, 16), 16), ((2, 5, 7, 9), 2), ((1, 1, 1, 1), 0), ((2, 3, 6, 5), 1)])
def test_option_a_shortcut_forward_backward_bit_exact(shape, pad):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).to(dev()).requires_grad_(True)
    y = PKG.ops.shortcut_a(x, pad)
    want = F.pad(x[:, :, ::2, ::2], (0, 0, 0, 0, pad, pad), "constant", 0.0)
    assert torch.equal(y, want)
    dy = torch.randn(y.shape, generator=g).to(dev())
    (gx,) = torch.autograd.grad(y, x, dy)
    (wx,) = torch.autograd.grad(want, x, dy)
    assert torch.equal(gx, wx)


@pytest.mark.parametrize("batch,inf,outf,bias", [(256, 64, 100, True), (128, 64, 10, True), (7, 5, 3, False), (1500, 300, 17, True), (1, 1, 1, True)])
def test_linear_weight_gradient_vs_fp64(batch, inf, outf, bias):
    g = torch.Generator().manual_seed(batch + inf)
    x = torch.randn(batch, inf, generator=g).to(dev()).requires_grad_(True)
    w = torch.randn(outf, inf, generator=g).to(dev()).requires_grad_(True)
    b = torch.randn(outf, generator=g).to(dev()).requires_grad_(True) if bias else None
    dy = torch.randn(batch, outf, generator=g).to(dev())
    y = PKG.ops.linear(x, w, b)
    assert torch.equal(y, F.linear(x, w, b))
    grads = torch.autograd.grad(y, [x, w] + ([b] if bias else []), dy)
    xd, wd, dyd = x.detach().double(), w.detach().double(), dy.double()
    torch.testing.assert_close(grads[0].double(), dyd @ wd, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(grads[1].double(), dyd.t() @ xd, rtol=1e-5, atol=1e-4)
    if bias:
        torch.testing.assert_close(grads[2].double(), dyd.sum(0), rtol=1e-5, atol=1e-4)
    # deterministic: a second evaluation is bit-identical
    again = torch.autograd.grad(PKG.ops.linear(x, w, b), [w], dy)[0]
    assert torch.equal(again, grads[1])


def test_device_prefetcher_delivers_every_batch_in_order():
    """prefetch.DevicePrefetcher: batch i+1 is copied on a side stream while batch i is consumed; contents, order and None
    entries must come through unchanged, also when the consumer keeps the GPU busy between batches."""
    g = torch.Generator().manual_seed(5)
    host = [(torch.rand(64, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 100, (64,), generator=g).pin_memory(), None)
            for _ in range(7)]
    busy = torch.randn(2048, 2048, device=dev())
    seen = []
    for x, y, u in PKG.prefetch.DevicePrefetcher(host, dev()):
        assert u is None and x.is_cuda and y.is_cuda
        busy = busy @ busy * 1e-3                                    # work on the consumer's stream
        seen.append((x.clone(), y.clone()))
    torch.cuda.synchronize()
    assert len(seen) == len(host)
    for (x, y), (hx, hy, _) in zip(seen, host):
        assert torch.equal(x.cpu(), hx) and torch.equal(y.cpu(), hy)
