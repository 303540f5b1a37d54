import importlib
import os

import numpy as np
import torch

PKG = importlib.import_module("cv_a-fan_b200")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev():
    return torch.device("cuda:0")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bitwise(a, b, what=""):
    """Bit-exact except that NaN payloads are not compared (NaN-ness must match)."""
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float32)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs"
    bad = (bits(a) != bits(b)) & ~na
    assert not bad.any(), f"{what}: {int(bad.sum())} of {a.size} elements differ bitwise"


def ulp_diff(a, b):
    """Max distance in units of last place between two finite float32 arrays."""
    ia = bits(a).astype(np.int64)
    ib = bits(b).astype(np.int64)
    ia = np.where(ia & 0x80000000, 0x80000000 - ia, ia)
    ib = np.where(ib & 0x80000000, 0x80000000 - ib, ib)
    return int(np.abs(ia - ib).max()) if ia.size else 0


class InjectGrad(torch.autograd.Function):
    """Scalar 'loss' whose gradient w.r.t. x is the prescribed tensor g (see oracle/gen_golden.py)."""

    @staticmethod
    def forward(ctx, x, g):
        ctx.save_for_backward(g)
        return x.new_zeros(())

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return g.clone(), None


class InjectingModel:
    def __init__(self, grads):
        self.grads, self.calls = grads, 0

    def _next(self, x_adv):
        g = self.grads[self.calls]
        self.calls += 1
        return InjectGrad.apply(x_adv, g)

    def __call__(self, x, end_point=None, start_point=None):
        return self._next(x["adv"] if isinstance(x, dict) else x)

    def train(self):
        return self

    def forward(self, inputs, bb, lb):
        out = self._next(inputs["adv"])
        z = out * 0
        return out, z, z, z


def load_pgd_cases():
    z = np.load(os.path.join(GOLDEN, "pgd_linf.npz"))
    for i in range(int(z["n_cases"])):
        k = f"case{i}"
        gamma, eps, steps, randinit, clip, flavour = z[k + "_meta"]
        yield dict(name=k, gamma=float(gamma), eps=float(eps), steps=int(steps), randinit=bool(randinit),
                   clip=bool(clip), flavour=int(flavour), x=z[k + "_x"], u=z[k + "_u"], grads=z[k + "_grads"],
                   states=z[k + "_states"], out=z[k + "_out"])


def feature_like(shape, gen):
    return torch.relu(1.5 * torch.randn(shape, generator=gen))


def grad_like(shape, gen):
    g = 1e-3 * torch.randn(shape, generator=gen)
    g[torch.rand(shape, generator=gen) < 0.05] = 0.0
    return g
