"""TEST-ONLY restatement of diffusers' StableDiffusion3PipelineOutput (a one-field BaseOutput)."""
from dataclasses import dataclass
from typing import Any

from ...utils import BaseOutput


@dataclass
class StableDiffusion3PipelineOutput(BaseOutput):
    images: Any = None
