"""Deterministic parameter fill shared by oracle/make_golden.py and the tests: lets a 67M-parameter
PipelineModel fixture be reproduced from a seed instead of being stored (TEST INFRASTRUCTURE)."""
import hashlib
import zlib

import torch


def deterministic_fill(model, seed):
    """Overwrite every parameter / buffer (in sorted key order; aliased tensors end with the fill of
    their last key) with seeded values of init-like magnitude."""
    sd = model.state_dict()
    with torch.no_grad():
        for key in sorted(sd):
            t = sd[key]
            if not t.is_floating_point() or key.endswith(".pe"):
                continue
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(key.encode()))
            if key.endswith("running_var"):
                v = torch.rand(t.shape, generator=g) + 0.5
            elif ".bns." in key and key.endswith("weight") or "norm" in key and key.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            elif t.dim() >= 2:
                v = torch.randn(t.shape, generator=g) / (t.shape[-1] ** 0.5)
            else:
                v = 0.05 * torch.randn(t.shape, generator=g)
            t.copy_(v)
    return model


def state_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
