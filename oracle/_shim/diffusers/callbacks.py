"""TEST-ONLY: the two callback classes custom_pipeline.py:8 imports for an isinstance check."""


class PipelineCallback:
    tensor_inputs = ["latents"]


class MultiPipelineCallbacks:
    tensor_inputs = ["latents"]
