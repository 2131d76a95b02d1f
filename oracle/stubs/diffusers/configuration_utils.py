import functools
import inspect
from types import SimpleNamespace


class ConfigMixin:
    pass


def register_to_config(init):
    """Run __init__, then expose the bound arguments as ``self.config`` (what diffusers does)."""

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        init(self, *args, **kwargs)
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        self.config = SimpleNamespace(**cfg)

    return wrapper
