class AutoencoderKL:
    @classmethod
    def from_pretrained(cls, *a, **k):
        raise NotImplementedError("diffusers stub: the VAE is outside the training hot path")
