import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)
    return load


def random_graphs(num_graphs, n_lo, n_hi, extra_per_node, seed, self_loops=True, isolated=False):
    """Random batched disjoint graphs (CPU): returns edge_index [2,E], batch [N]."""
    g = torch.Generator().manual_seed(seed)
    src, dst, batch, off = [], [], [], 0
    for b in range(num_graphs):
        n = int(torch.randint(n_lo, n_hi + 1, (1,), generator=g))
        if self_loops:
            src += list(range(off, off + n)); dst += list(range(off, off + n))
        m = int(extra_per_node * n)
        lo = 1 if (isolated and n > 1) else 0   # node `off` never receives a random edge
        s = torch.randint(0, n, (m,), generator=g) + off
        d = torch.randint(lo, n, (m,), generator=g) + off
        src += s.tolist(); dst += d.tolist()
        batch += [b] * n
        off += n
    ei = torch.tensor([src, dst], dtype=torch.long).reshape(2, -1)
    return ei, torch.tensor(batch, dtype=torch.long)
