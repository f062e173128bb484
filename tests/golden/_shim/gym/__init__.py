"""Minimal `gym` stand-in, used ONLY by tests/golden/make_golden.py.

The reference's `utils/__init__.py` imports gym transitively, so importing its
hot-path modules (learner, policies, dsgd, utils.noise_sources) needs *a* gym on
sys.path. gym is not installed in the build container and there is no network.
Nothing on the product path imports this.
"""
from . import spaces  # noqa: F401

_REGISTRY = {}


class Env(object):
    metadata = {}
    observation_space = None
    action_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def seed(self, seed=None):
        return [seed]


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    import importlib
    entry_point, reg_kwargs = _REGISTRY[id]
    mod_name, cls_name = entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    merged = dict(reg_kwargs.get("kwargs", {}))
    merged.update(kwargs)
    return cls(**merged)
