"""Deterministic, name-keyed parameter fill shared by make_golden.py (applied to the reference model) and the tests
(applied to the model under test), so that large parameter sets need not be stored in the fixtures."""
import zlib

import torch


def fill_by_name(module):
    with torch.no_grad():
        for name, v in module.state_dict().items():
            if not v.dtype.is_floating_point:
                continue
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
            r = torch.randn(v.shape, generator=g)
            if v.ndim >= 2:
                fan_in = v[0].numel()
                r = r * (1.0 / fan_in ** 0.5)
            elif name.endswith("weight"):
                r = 1.0 + 0.1 * r
            else:
                r = 0.02 * r
            v.copy_(r.to(v.dtype))
    return module


def grad_digest(g, full_below=4096, head=512):
    """(kind, array): the full gradient for small tensors, else [sum, l2, first `head` entries]."""
    flat = g.detach().reshape(-1).double()
    if flat.numel() <= full_below:
        return g.detach().numpy()
    return torch.cat([flat.sum()[None], flat.norm()[None], flat[:head]]).numpy()
