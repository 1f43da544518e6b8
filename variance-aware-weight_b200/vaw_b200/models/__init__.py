"""Denoisers behind the reference's constructors: DiT / U-ViT run on the B200 engines, the UNet is a PyTorch module."""


def model_table():
    """name -> constructor over the three families (the reference keeps DiT_models / UViT_models / UNet_models)."""
    from .dit import DiT_models
    from .unet import UNet_models
    from .uvit import UViT_models
    return {**DiT_models, **UViT_models, **UNet_models}
