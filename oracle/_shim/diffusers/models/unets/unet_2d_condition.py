class UNet2DConditionModel:  # imported by backbones/animatediff/models/sparse_controlnet.py (never constructed on the path)
    pass
