import functools
import inspect
from types import SimpleNamespace


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        if not hasattr(self, "_internal_dict"):
            self._internal_dict = _Config()
        self._internal_dict.update(kwargs)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for p in sig.parameters.values() if p.name != "self"]
        cfg = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty
               and p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)}
        names = [p.name for p in params if p.kind == p.POSITIONAL_OR_KEYWORD]
        for n, a in zip(names, args):
            cfg[n] = a
        cfg.update(kwargs)
        self.register_to_config(**cfg)
        init(self, *args, **kwargs)

    return inner


class FrozenDict(dict):
    pass
