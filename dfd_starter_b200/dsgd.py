"""`DSGD` with the reference's constructor and `adjust_lr` (dsgd/dynamic_sgd.py:7-51).

Inside `FiniteDifferences.step` the update theta += lr*sqrt(P)*lr_scale*g/||g|| runs
on the device (dfd_dsgd_step); this class carries the hyper-parameters and the
omega -> lr_scale map, and is recognised by the learner the way the reference
recognises its own class (finite_differences.py:22)."""
import numpy as np
import torch
from torch.optim import Optimizer


def affine_transform(value, from_min, from_max, to_min, to_max):
    """utils/math_helpers.py:137-144."""
    if from_max == from_min or to_max == to_min:
        return to_min
    mapped = (value - from_min) * (to_max - to_min) / (from_max - from_min)
    mapped += to_min
    return mapped


class DSGD(Optimizer):
    """Hyper-parameters of the normalised-gradient step theta -= lr * sqrt(P) * lr_scale * g / ||g||.  The attribute
    names (`lr`, `coef`, `lr_scale`, `min_scale`, `max_scale`, `steps`) are the ones the reference learner and drivers
    read (dynamic_sgd.py:7-16, finite_differences.py:51-52)."""

    def __init__(self, params, lr, min_scale=0.23, max_scale=1.0):
        super().__init__(params, {"lr": lr})
        self.lr, self.min_scale, self.max_scale = lr, min_scale, max_scale
        self.lr_scale = 1
        self.steps = 0
        self.coef = np.sqrt(sum(p.numel() for group in self.param_groups for p in group["params"]))

    @torch.no_grad()
    def step(self, closure=None):
        """Stand-alone optimizer step on whatever device the parameters live on
        (dynamic_sgd.py:18-39).  The learner does not come through here."""
        grads = [p.grad.reshape(-1) for g in self.param_groups for p in g["params"] if p.grad is not None]
        flat_grad = torch.cat(grads) if grads else torch.zeros(0)
        norm = flat_grad.norm().item()
        assert norm > 0, "DSGD ENCOUNTERED GRADIENT WITH NORM OF ZERO"
        coef = self.lr * self.coef * self.lr_scale / norm
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    p.sub_(coef * p.grad)
        self.steps += 1

    def adjust_lr(self, omega):
        self.lr_scale = affine_transform(omega.omega, omega.min_omega, omega.max_omega, self.min_scale, self.max_scale)


def is_dsgd(opt):
    """The reference tests `type(opt) == DSGD` against its own class; accept that
    class too (same name and fields) so a reference DSGD object can be passed in."""
    return type(opt) is DSGD or (type(opt).__name__ == "DSGD" and all(
        hasattr(opt, a) for a in ("lr", "lr_scale", "min_scale", "max_scale", "adjust_lr")))
