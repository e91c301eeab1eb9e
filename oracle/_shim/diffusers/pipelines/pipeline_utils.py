import contextlib

import torch


class DiffusionPipeline:
    """Only what the reference pipeline uses: register_modules, .device, .progress_bar, numpy_to_pil."""

    def register_modules(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self):
        return torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        class _Bar:
            def update(self, n=1):
                pass
        yield _Bar()
