import os
import sys

import numpy as np
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


def make_case(views, wh, seed_batch=0, seed_scene=1):
    """Seeded synthetic inputs shared by the oracle, the golden generator and the CUDA path."""
    from uforecon_b200 import checkpoint, synthetic
    sd = checkpoint.synthetic_state_dict(0)
    batch = synthetic.make_batch(views, wh, seed=seed_batch)
    scene = synthetic.make_scene(batch, seed=seed_scene)
    batch["depth_info"] = scene["depth_info"]
    return batch, scene, sd


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    return {k: g[k] for k in g.files}


def rel_err(a, b):
    """max |a-b| / max|b| : the 'relative' of the 1e-5 parity bar (scale of the reference tensor)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
