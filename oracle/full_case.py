"""Recipe + seed-regenerated inputs of the FULL-SIZE training golden (tests/golden/cls_train_full.npz).

TEST INFRASTRUCTURE (like everything under oracle/): imported by oracle/gen_golden.py, which executes the unmodified
reference loop on these inputs, and by the parity tests, which feed the same inputs to the CPU port and to the CUDA path.
"""
import torch

FULL_RECIPE = dict(num_blocks=[9, 9, 9], num_classes=100, perturb_idx=13, steps=5, gamma=0.5, eps=2.0, randinit=True,
                   clip=True, batch=128, iters=2, epoch=1, weight_seed=3, data_seed=11, noise_seed=3, sub=8)


def full_case_inputs(recipe=FULL_RECIPE):
    """Inputs of the full-size golden, REGENERATED from seeds (CPU generator streams are deterministic): the vectors
    themselves (2 x 128 x 16 x 32 x 32 noise = 16.8 MB) are too large to commit.  Used by the generator and by the tests."""
    gen = torch.Generator().manual_seed(recipe["data_seed"])
    bs, iters = recipe["batch"], recipe["iters"]
    images = [torch.rand(bs, 3, 32, 32, generator=gen) for _ in range(iters)]
    targets = [torch.randint(0, recipe["num_classes"], (bs,), generator=gen) for _ in range(iters)]
    ngen = torch.Generator().manual_seed(recipe["noise_seed"])
    noises = [torch.rand(bs, 16, 32, 32, generator=ngen) for _ in range(iters)]     # idx 13 of [9,9,9]: stage-1 output
    return images, targets, noises
