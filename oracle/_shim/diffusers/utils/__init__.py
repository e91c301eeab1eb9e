import logging as _logging
from collections import OrderedDict
from dataclasses import fields


class BaseOutput(OrderedDict):
    """Dataclass-style output with both attribute and item access."""

    def __post_init__(self):
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                super().__setitem__(f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return tuple(self.values())[k]


class _Logging:
    @staticmethod
    def get_logger(name):
        return _logging.getLogger(name)


logging = _Logging()


def deprecate(*args, **kwargs):
    return None


def is_accelerate_available():
    return False


def export_to_video(*args, **kwargs):  # imported by inversion_tools/flow_inversion.py, unused on the tested path
    raise NotImplementedError
