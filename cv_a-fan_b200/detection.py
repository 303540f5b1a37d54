"""Drop-in for Detection/attack_algo.py's hot-path functions (Faster R-CNN flavour).

    PGD(x, image_batch, y, model, steps, eps, gamma, idx, randinit, clip)   (Detection/attack_algo.py:48-74)
    compute_loss(l1, l2, l3, l4)                                            (:21-27)
    mix_feature / get_sample_points                                         (:254-265 / :236-245)
The model contract is the reference's: model.train().forward({'x','adv','out_idx','flag'}, bb, lb) -> 4 losses.
"""
from .attack_algo import linfball_proj, l2ball_proj, pgd_loop  # noqa: F401
from .segmentation import get_sample_points, mix_feature, sat_sample_points  # noqa: F401  (identical maths in both reference files)


def compute_loss(loss1, loss2, loss3, loss4):
    return loss1.mean() + loss2.mean() + loss3.mean() + loss4.mean()


def PGD(x, image_batch, y=None, model=None, steps=3, eps=None, gamma=None, idx=1, randinit=False, clip=False,
        **extras):
    def tail_loss(x_adv):
        inputs = {"x": image_batch, "adv": x_adv, "out_idx": idx, "flag": "tail"}
        a_obj, a_trf, p_cls, p_trf = model.train().forward(inputs, y["bb"], y["lb"])
        return compute_loss(a_obj, a_trf, p_cls, p_trf)

    return pgd_loop(x, tail_loss, steps, gamma, eps, randinit, clip, **extras)
