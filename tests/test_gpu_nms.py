"""-m gpu: device-side NMS vs the reference's own golden vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from tests.util import GOLDEN, PKG, dev

pytestmark = pytest.mark.gpu


def test_reference_unit_tests_pass():
    """The four cases of Detection/test/nms/test_nms.py, through the drop-in `detection.nms`."""
    nms = PKG.detection.nms
    assert len(nms(torch.tensor([], dtype=torch.float).to(dev()), torch.tensor([], dtype=torch.float).to(dev()), 0.7)) == 0
    assert nms(torch.tensor([[5, 5, 10, 10]], dtype=torch.float).to(dev()), torch.tensor([0.8]).to(dev()), 0.7).tolist() == [0]
    b = torch.tensor([[5, 5, 10, 10], [5, 5, 10, 10], [5, 5, 30, 30]], dtype=torch.float).to(dev())
    assert nms(b, torch.tensor([0.6, 0.9, 0.4]).to(dev()), 0.7).tolist() == [1, 2]
    z = np.load(os.path.join(GOLDEN, "nms_large.npz"))
    det = torch.from_numpy(z["input"]).float().to(dev())
    kept = nms(det[:, :4], det[:, 4], 0.7)
    assert len(kept) == 1934
    assert sorted(kept.tolist()) == sorted(z["output"].tolist())
    flags, count = PKG.ops.nms_flags(det[:, :4].contiguous(), det[:, 4].contiguous(), 0.7)
    assert int(count) == 1934 and int(flags.sum()) == 1934


@pytest.mark.parametrize("n,thr", [(1, 0.5), (63, 0.5), (64, 0.3), (65, 0.7), (1000, 0.5), (12000, 0.7)])
def test_nms_vs_oracle_random(n, thr):
    g = torch.Generator().manual_seed(n)
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    boxes = torch.cat([xy, xy + wh], 1).round()                   # integer coordinates -> many exact-threshold-free IoUs
    scores = torch.rand(n, generator=g)
    kept = PKG.detection.nms(boxes.to(dev()), scores.to(dev()), thr).cpu().numpy()
    assert np.array_equal(kept, orc.nms(boxes.numpy(), scores.numpy(), thr, strict_gt=True))


@pytest.mark.parametrize("images,n,thr,max_keep", [(1, 1, 0.5, 4), (3, 65, 0.7, 1000), (4, 1000, 0.5, 37), (8, 12000, 0.7, 2000),
                                                   (2, 3000, 0.7, 64), (2, 500, 0.3, 128)])
def test_batched_proposal_nms_vs_oracle(images, n, thr, max_keep):
    """afan_nms_batched_f32: one CTA per image, early exit at max_keep, compacted zero-padded output, flags by rank --
    against the C restatement applied per image followed by the reference's `[:post_nms_top_n]` slice."""
    g = torch.Generator().manual_seed(images * 131 + n)
    xy = torch.rand(images, n, 2, generator=g) * 300
    wh = torch.rand(images, n, 2, generator=g) * 120 + 4
    boxes = torch.cat([xy, xy + wh], 2).round()
    kept, counts, flags = PKG.ops.nms_batched(boxes.to(dev()).contiguous(), thr, max_keep, want_flags=True)
    kept, counts, flags = kept.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy()
    for i in range(images):
        scores = -np.arange(n, dtype=np.float32)                   # already ranked
        want = orc.nms(boxes[i].numpy(), scores, thr, strict_gt=True)[:max_keep]
        assert counts[i] == len(want)
        assert np.array_equal(np.nonzero(flags[i])[0], want)
        assert np.array_equal(kept[i, :len(want)], boxes[i].numpy()[want])
        assert not kept[i, len(want):].any()
