"""TEST INFRASTRUCTURE - deterministic, name-keyed parameter fill shared by the golden-fixture
generator (reference modules) and the parity tests (product modules).  Because both sides expose the
same state_dict keys, filling by key gives bit-identical weights without shipping 200 M parameters."""
import math
import zlib

import torch


@torch.no_grad()
def det_fill_(module, seed=0):
    """Fill every floating tensor of `module.state_dict()` (or of a plain dict name -> tensor)."""
    sd = module if isinstance(module, dict) else module.state_dict()
    for name in sorted(sd.keys()):
        t = sd[name]
        if not t.is_floating_point():
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "running_var":
            v = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif "sampling_offsets.bias" in name:
            v = torch.randn(shape, generator=g) * 1.5                 # offsets of a few cells
        elif "gamma_" in name:
            v = 0.25 + 0.05 * torch.randn(shape, generator=g)         # ALIF gates
        elif t.dim() >= 2:
            fan_in = max(1, t[0].numel())
            v = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
        elif leaf == "weight":                                         # norm scales
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                                          # biases, running_mean, ...
            v = 0.05 * torch.randn(shape, generator=g)
        t.copy_(v.to(t.dtype))
    return module
